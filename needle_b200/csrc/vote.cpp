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

// ComparatorHeapEntry (comparator.rs:20-35).  The derived Ord is lexicographic over
// the fields in declaration order: score, src_longest_run, dst_longest_run,
// src_match_hash, dst_match_hash, the four is_* flags, the two hash durations.
// Within one heap (one table of one pair) the flags and hash durations are the
// same for every entry, so they never decide a comparison.
struct Entry {
    uint64_t score, src_start, src_end, dst_start, dst_end;
    uint32_t src_hash, dst_hash;
};

inline bool entry_greater(const Entry &a, const Entry &b) {
    if (a.score != b.score) return a.score > b.score;
    if (a.src_start != b.src_start) return a.src_start > b.src_start;
    if (a.src_end != b.src_end) return a.src_end > b.src_end;
    if (a.dst_start != b.dst_start) return a.dst_start > b.dst_start;
    if (a.dst_end != b.dst_end) return a.dst_end > b.dst_end;
    if (a.src_hash != b.src_hash) return a.src_hash > b.src_hash;
    return a.dst_hash > b.dst_hash;
}

// BinaryHeap::push on heap[base..): append, then sift the new element up while it
// is greater than its parent.
void heap_push(std::vector<Entry> &heap, size_t base, const Entry &e) {
    heap.push_back(e);
    size_t pos = heap.size() - 1 - base;
    while (pos > 0) {
        const size_t parent = (pos - 1) / 2;
        if (!entry_greater(e, heap[base + parent])) break;
        heap[base + pos] = heap[base + parent];
        pos = parent;
    }
    heap[base + pos] = e;
}

struct Candidate {
    uint64_t start, end, hash_duration;
    uint32_t match_hash;
    bool is_opening;
};

}  // namespace

int vote_impl(const uint64_t *hash_duration_ns, uint32_t n_videos, const uint32_t (*pairs)[2],
              uint64_t n_pairs, const nb200_match_params *params, const nb200_run *runs,
              uint64_t n_runs, const uint8_t *video_mask, nb200_search_result *results) {
    std::vector<uint32_t> default_flat;
    if (!pairs) {
        default_pairs(n_videos, default_flat);
        pairs = reinterpret_cast<const uint32_t(*)[2]>(default_flat.data());
        n_pairs = default_flat.size() / 2;
    }
    std::memset(results, 0, sizeof(nb200_search_result) * n_videos);

    // Replay every table's heap from its runs (already in push order).  All heaps live
    // back to back in one array; group g = 2 * pair + is_ending covers
    // heap[group_begin[g] .. group_begin[g + 1]).
    std::vector<Entry> heap;
    heap.reserve(n_runs);
    std::vector<uint64_t> group_begin(2 * n_pairs + 1, 0);
    {
        uint64_t g_prev = 0;
        for (uint64_t r = 0; r < n_runs; r++) {
            const nb200_run &run = runs[r];
            if (run.pair >= n_pairs || run.is_ending > 1) return NB200_ERR_INVALID_ARGUMENT;
            if (pairs[run.pair][0] >= n_videos || pairs[run.pair][1] >= n_videos) return NB200_ERR_INVALID_ARGUMENT;
            if (r > 0) {   // (pair, is_ending, i desc, j desc)
                const nb200_run &q = runs[r - 1];
                auto key = [](const nb200_run &x) {
                    return std::make_tuple(x.pair, x.is_ending, ~x.i_end, ~x.j_end);
                };
                if (!(key(q) < key(run))) return NB200_ERR_INVALID_ARGUMENT;
            }
            const uint64_t g = 2ull * run.pair + run.is_ending;
            for (uint64_t k = g_prev + 1; k <= g; k++) group_begin[k] = heap.size();
            g_prev = g;
            const Entry e = {run.len, run.src_start_ns, run.src_end_ns, run.dst_start_ns, run.dst_end_ns,
                             run.src_simhash, run.dst_simhash};
            heap_push(heap, (size_t)group_begin[g], e);
        }
        for (uint64_t k = g_prev + 1; k <= 2 * n_pairs; k++) group_begin[k] = heap.size();
    }

    // info_map: for every video the non-empty pair infos it takes part in, in pair
    // order, with whether it is the source (:562, :580-588).  Counting sort by video.
    std::vector<uint64_t> vid_begin(n_videos + 1, 0);
    auto pair_nonempty = [&](uint64_t k) { return group_begin[2 * k + 2] > group_begin[2 * k]; };
    for (uint64_t k = 0; k < n_pairs; k++)
        if (pair_nonempty(k)) {
            vid_begin[pairs[k][0] + 1]++;
            vid_begin[pairs[k][1] + 1]++;
        }
    for (uint32_t v = 0; v < n_videos; v++) vid_begin[v + 1] += vid_begin[v];
    std::vector<uint64_t> vid_pairs(vid_begin[n_videos]);   // 2 * pair + is_source
    {
        std::vector<uint64_t> cursor(vid_begin.begin(), vid_begin.end() - 1);
        for (uint64_t k = 0; k < n_pairs; k++)
            if (pair_nonempty(k)) {
                // a pair (v, v) cannot come from the reference's pair list, but keep its order: source first
                vid_pairs[cursor[pairs[k][0]]++] = 2 * k + 1;
                vid_pairs[cursor[pairs[k][1]]++] = 2 * k;
            }
    }

    int status = NB200_OK;
    const uint32_t T = params->hash_match_threshold;
    const uint32_t bias = T + T / 2;
    std::vector<Candidate> cand;
    std::vector<uint32_t> cluster, uniq, mult, ucount;
    for (uint32_t v = 0; v < n_videos; v++) {
        if (video_mask && !video_mask[v]) continue;        // another rank votes for this video
        if (vid_begin[v + 1] == vid_begin[v]) continue;   // find_best_match -> None
        cand.clear();
        for (uint64_t m = vid_begin[v]; m < vid_begin[v + 1]; m++) {
            const uint64_t k = vid_pairs[m] >> 1;
            const bool is_source = vid_pairs[m] & 1;
            const uint64_t hd = hash_duration_ns[is_source ? pairs[k][0] : pairs[k][1]];
            for (int e = 0; e < 2; e++)   // openings, then endings (:413-431)
                for (uint64_t x = group_begin[2 * k + e]; x < group_begin[2 * k + e + 1]; x++) {
                    const Entry &en = heap[x];
                    Candidate c;
                    c.start = is_source ? en.src_start : en.dst_start;
                    c.end = is_source ? en.src_end : en.dst_end;
                    c.match_hash = is_source ? en.src_hash : en.dst_hash;
                    c.hash_duration = hd;
                    c.is_opening = e == 0;
                    cand.push_back(c);
                }
        }
        const size_t nc = cand.size();
        // |distinct_matches[i]| = #{ j : popcount(h_i ^ h_j) < T + T/2 }; i is a key iff that is > 0
        // (the relation is symmetric: visit each unordered pair once)
        // The same intro seen from many pairs gives few distinct signatures: count over the
        // distinct values with their multiplicities (u^2 instead of c^2 distances).
        uniq.clear();
        for (size_t i = 0; i < nc; i++) uniq.push_back(cand[i].match_hash);
        std::sort(uniq.begin(), uniq.end());
        mult.clear();
        {
            size_t w = 0;
            for (size_t i = 0; i < uniq.size(); i++) {
                if (w > 0 && uniq[w - 1] == uniq[i]) {
                    mult[w - 1]++;
                } else {
                    uniq[w++] = uniq[i];
                    mult.push_back(1);
                }
            }
            uniq.resize(w);
        }
        const size_t nu = uniq.size();
        ucount.assign(nu, 0);
        for (size_t a = 0; a < nu; a++) {
            if (bias > 0) ucount[a] += mult[a];          // distance 0 to every copy of itself (including itself)
            const uint32_t ha = uniq[a], ma = mult[a];
            uint32_t mine = 0;
            for (size_t b = a + 1; b < nu; b++) {      // branch-free: the outcome is data-random
                const uint32_t hit = (uint32_t)__builtin_popcount(ha ^ uniq[b]) < bias ? 1u : 0u;
                mine += hit * mult[b];
                ucount[b] += hit * ma;
            }
            ucount[a] += mine;
        }
        cluster.resize(nc);
        for (size_t i = 0; i < nc; i++)
            cluster[i] = ucount[std::lower_bound(uniq.begin(), uniq.end(), cand[i].match_hash) - uniq.begin()];

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
