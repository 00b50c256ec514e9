// vote.cpp -- host tail of the match path: what Comparator::run_with_frame_hashes
// does after the per-pair searches (needle/src/audio/comparator.rs:580-626),
// reproduced exactly so that the final SearchResult intervals are bit-identical:
//
//   * the pair list order                         (:534-545)
//   * each table's runs pushed into a BinaryHeap in (i desc, j desc) order and
//     read back as the heap's internal array       (:191-192, :231-249)
//   * find_opening_and_ending's entry lists        (:283-300)
//   * info_map and find_best_match                 (:580-588, :405-515)
//
// Pure C++ (no CUDA): also used by rank 0 of a multi-GPU job on gathered runs.
// Every run record carries its four timestamps, so nothing else is read.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <tuple>
#include <vector>

#include "common.h"

namespace nb200 {

void default_pairs(uint32_t n_videos, std::vector<uint32_t> &flat) {
    // for i: for j: skip i == j and already-processed j  =>  all (i < j), i-major
    flat.clear();
    std::vector<char> processed(n_videos, 0);
    for (uint32_t i = 0; i < n_videos; i++) {
        for (uint32_t j = 0; j < n_videos; j++) {
            if (i == j || processed[j]) continue;
            flat.push_back(i);
            flat.push_back(j);
        }
        processed[i] = 1;
    }
}

namespace {

// Rust Duration::as_secs_f32: (secs as f32) + (nanos as f32) / 1e9
float as_secs_f32(uint64_t ns) {
    volatile float s = (float)(ns / 1000000000ull);
    volatile float f = (float)(uint32_t)(ns % 1000000000ull) / 1000000000.0f;
    return s + f;
}

// ComparatorHeapEntry, fields in declaration order so that std::tuple's
// lexicographic operator< equals the derived Ord (comparator.rs:20-35).
using Entry = std::tuple<uint64_t,                       // score
                         uint64_t, uint64_t,             // src_longest_run
                         uint64_t, uint64_t,             // dst_longest_run
                         uint32_t, uint32_t,             // src/dst_match_hash
                         bool, bool, bool, bool,         // is_src_opening, is_src_ending, is_dst_..
                         uint64_t, uint64_t>;            // src/dst_hash_duration

// BinaryHeap::push: append then sift the new element up while it is greater
// than its parent.
void heap_push(std::vector<Entry> &heap, const Entry &e) {
    heap.push_back(e);
    size_t pos = heap.size() - 1;
    while (pos > 0) {
        const size_t parent = (pos - 1) / 2;
        if (!(e > heap[parent])) break;
        heap[pos] = heap[parent];
        pos = parent;
    }
    heap[pos] = e;
}

struct Candidate {
    uint64_t start, end, hash_duration;
    uint32_t match_hash;
    bool is_opening;
};

struct PairInfo {            // OpeningAndEndingInfo; src_* and dst_* lists hold the same entries
    std::vector<Entry> openings, endings;
    bool empty() const { return openings.empty() && endings.empty(); }
};

}  // namespace

int vote_impl(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
              uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
              uint64_t n_runs, nb200_search_result *results) {
    std::vector<uint32_t> default_flat;
    if (!pairs) {
        default_pairs(n_videos, default_flat);
        pairs = reinterpret_cast<const uint32_t(*)[2]>(default_flat.data());
        n_pairs = default_flat.size() / 2;
    }
    std::memset(results, 0, sizeof(nb200_search_result) * n_videos);

    // rebuild every pair's heap arrays from the runs (already in push order)
    std::vector<PairInfo> infos(n_pairs);
    for (uint64_t r = 0; r < n_runs; r++) {
        const nb200_run &run = runs[r];
        if (run.pair >= n_pairs) return NB200_ERR_INVALID_ARGUMENT;
        if (r > 0) {   // (pair, is_ending, i desc, j desc)
            const nb200_run &q = runs[r - 1];
            auto key = [](const nb200_run &x) {
                return std::make_tuple(x.pair, x.is_ending, ~x.i_end, ~x.j_end);
            };
            if (!(key(q) < key(run))) return NB200_ERR_INVALID_ARGUMENT;
        }
        const uint32_t src = pairs[run.pair][0], dst = pairs[run.pair][1];
        if (src >= n_videos || dst >= n_videos) return NB200_ERR_INVALID_ARGUMENT;
        const bool is_opening = !run.is_ending;
        Entry en(run.len, run.src_start_ns, run.src_end_ns, run.dst_start_ns, run.dst_end_ns,
                 run.src_simhash, run.dst_simhash, is_opening, !is_opening, is_opening, !is_opening,
                 hash_duration_ns[src], hash_duration_ns[dst]);
        heap_push(run.is_ending ? infos[run.pair].endings : infos[run.pair].openings, en);
    }

    // info_map: only non-empty infos, in pair order (:562, :580-588)
    std::vector<std::vector<std::pair<const PairInfo *, bool>>> info_map(n_videos);
    for (uint64_t k = 0; k < n_pairs; k++) {
        if (infos[k].empty()) continue;
        info_map[pairs[k][0]].push_back({&infos[k], true});
        info_map[pairs[k][1]].push_back({&infos[k], false});
    }

    int status = NB200_OK;
    const uint32_t T = params->hash_match_threshold;
    const uint32_t bias = T + T / 2;
    for (uint32_t v = 0; v < n_videos; v++) {
        const auto &matches = info_map[v];
        if (matches.empty()) continue;   // find_best_match -> None
        std::vector<Candidate> cand;
        for (const auto &mi : matches) {
            const PairInfo *m = mi.first;
            const bool is_source = mi.second;
            for (int pass = 0; pass < 2; pass++) {
                for (const Entry &e : (pass == 0 ? m->openings : m->endings)) {
                    Candidate c;
                    if (is_source) {
                        c.start = std::get<1>(e);
                        c.end = std::get<2>(e);
                        c.hash_duration = std::get<11>(e);
                        c.match_hash = std::get<5>(e);
                    } else {
                        c.start = std::get<3>(e);
                        c.end = std::get<4>(e);
                        c.hash_duration = std::get<12>(e);
                        c.match_hash = std::get<6>(e);
                    }
                    c.is_opening = pass == 0;
                    cand.push_back(c);
                }
            }
        }
        const size_t nc = cand.size();
        // |distinct_matches[i]| = #{ j : popcount(h_i ^ h_j) < T + T/2 }; i is a key iff that is > 0
        // (the relation is symmetric: visit each unordered pair once)
        std::vector<uint32_t> cluster(nc, 0);
        for (size_t i = 0; i < nc; i++) {
            const uint32_t hi = cand[i].match_hash;
            if (0u < bias) cluster[i]++;   // i with itself: distance 0
            for (size_t j = i + 1; j < nc; j++)
                if ((uint32_t)__builtin_popcount(hi ^ cand[j].match_hash) < bias) {
                    cluster[i]++;
                    cluster[j]++;
                }
        }

        nb200_search_result &res = results[v];
        res.present = 1;
        for (int want_opening = 1; want_opening >= 0; want_opening--) {
            if (!want_opening && !params->include_endings) break;
            bool have = false;
            float best_score = 0.f;
            size_t best_k = 0;
            for (size_t k = 0; k < nc; k++) {
                if (cluster[k] == 0 || cand[k].is_opening != (want_opening != 0)) continue;
                if (cand[k].end < cand[k].start) {
                    status = NB200_ERR_DURATION_UNDERFLOW;
                    continue;
                }
                volatile float a = (float)(int64_t)cluster[k] * 0.3f;
                volatile float b = as_secs_f32(cand[k].end - cand[k].start) * 0.7f;
                volatile float sum = a + b;
                const float score = -sum;
                // sort ascending by (score, k), take the first  ==  argmin with ties to smaller k
                if (!have || score < best_score) {
                    have = true;
                    best_score = score;
                    best_k = k;
                }
            }
            if (!have) continue;
            const Candidate &w = cand[best_k];
            const uint64_t sub = params->time_padding_ns + w.hash_duration;
            if (w.end < sub) {
                status = NB200_ERR_DURATION_UNDERFLOW;   // end - padding - hash_duration panics
                continue;
            }
            if (want_opening) {
                res.has_opening = 1;
                res.opening_start_ns = w.start + params->time_padding_ns;
                res.opening_end_ns = w.end - sub;
            } else {
                res.has_ending = 1;
                res.ending_start_ns = w.start + params->time_padding_ns;
                res.ending_end_ns = w.end - sub;
            }
        }
    }
    return status;
}

}  // namespace nb200
