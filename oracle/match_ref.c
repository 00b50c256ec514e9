/*
 * match_ref.c -- CPU restatement of needle's Comparator (TEST INFRASTRUCTURE,
 * see needle_oracle.h; parity unpinned -- the reference has no tests for it).
 *
 * Follows /root/reference/needle/src/audio/comparator.rs:
 *   orc_simhash32                  <- :149-153 (+ chromaprint-rust 0.1.3 simhash32)
 *   orc_longest_common_hash_match  <- :157-250
 *   find_opening_and_ending        <- :252-308
 *   find_best_match                <- :405-515
 *   orc_run_with_frame_hashes      <- :524-629
 * and Rust std semantics for Duration (f32 conversions) and BinaryHeap::push.
 *
 * Build without -ffast-math and with -ffp-contract=off: the f32 score
 * arithmetic of find_best_match must round exactly like Rust's.
 */
#include "needle_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Duration */

uint64_t orc_duration_from_secs_f32(float secs) {
    if (!(secs >= 0.0f) || isinf(secs)) return UINT64_MAX;
    /* f32 has a 24-bit significand and 1e9 = 2^9 * 5^9 with 5^9 < 2^21, so the
     * double product is exact; nearbyint under the default rounding mode is
     * round-half-to-even, which is what Duration::from_secs_f32 does. */
    double ns = nearbyint((double)secs * 1e9);
    if (ns >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)ns;
}

float orc_duration_as_secs_f32(uint64_t ns) {
    uint64_t secs = ns / 1000000000ull;
    uint32_t nanos = (uint32_t)(ns % 1000000000ull);
    volatile float a = (float)secs;
    volatile float b = (float)nanos / 1000000000.0f;
    return a + b;
}

uint64_t orc_duration_mul_f32(uint64_t ns, float rhs) {
    volatile float s = orc_duration_as_secs_f32(ns);
    volatile float p = rhs * s;
    return orc_duration_from_secs_f32(p);
}

uint64_t orc_hash_timestamp(uint64_t delay_ns, uint64_t item_ns, uint32_t raw_index,
                            uint64_t seek_to_ns) {
    return delay_ns + orc_duration_mul_f32(item_ns, (float)raw_index) + seek_to_ns;
}

/* ------------------------------------------------------------------- simhash */

uint32_t orc_simhash32(const uint32_t *hashes, size_t n) {
    int v[32];
    memset(v, 0, sizeof v);
    for (size_t k = 0; k < n; k++) {
        uint32_t h = hashes[k];
        for (int b = 0; b < 32; b++) v[b] += (h & (1u << b)) ? 1 : -1;
    }
    uint32_t out = 0;
    for (int b = 0; b < 32; b++)
        if (v[b] > 0) out |= 1u << b;
    return out;
}

/* --------------------------------------------------- ComparatorHeapEntry Ord */

/* #[derive(PartialOrd, Ord)] on comparator.rs:20-35: lexicographic over the
 * fields in declaration order.  Booleans: false < true. */
static int entry_cmp(const orc_entry *a, const orc_entry *b) {
#define CMP(f)                     \
    if (a->f != b->f) return a->f < b->f ? -1 : 1;
    CMP(score)
    CMP(src_start_ns) CMP(src_end_ns)
    CMP(dst_start_ns) CMP(dst_end_ns)
    CMP(src_match_hash) CMP(dst_match_hash)
    /* is_src_opening, is_src_ending, is_dst_opening, is_dst_ending */
    {
        int ao = !a->is_ending, bo = !b->is_ending;
        if (ao != bo) return ao < bo ? -1 : 1;           /* is_src_opening */
        if (a->is_ending != b->is_ending) return a->is_ending < b->is_ending ? -1 : 1;
    }
    CMP(src_hash_duration_ns) CMP(dst_hash_duration_ns)
#undef CMP
    return 0;
}

typedef struct {
    orc_entry *data;
    size_t len, cap;
} heap_t;

/* std::collections::BinaryHeap::push = Vec::push + sift_up(0, old_len):
 * move the hole up while element > parent. */
static void heap_push(heap_t *h, const orc_entry *e) {
    if (h->len == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 8;
        h->data = (orc_entry *)realloc(h->data, h->cap * sizeof(orc_entry));
    }
    size_t pos = h->len++;
    while (pos > 0) {
        size_t parent = (pos - 1) / 2;
        if (entry_cmp(e, &h->data[parent]) <= 0) break;
        h->data[pos] = h->data[parent];
        pos = parent;
    }
    h->data[pos] = *e;
}

/* ----------------------------------------------- longest_common_hash_match */

int64_t orc_longest_common_hash_match(const uint32_t *src_hash, const uint64_t *src_ts, size_t n,
                                      const uint32_t *dst_hash, const uint64_t *dst_ts, size_t m,
                                      uint32_t threshold, uint64_t min_opening_ns,
                                      uint64_t min_ending_ns, uint64_t src_hash_duration_ns,
                                      uint64_t dst_hash_duration_ns, int is_opening,
                                      orc_entry **out) {
    *out = NULL;
    if (n == 0 || m == 0) return 0; /* :165-167 */

    int is_ending = !is_opening;
    heap_t heap = {0};
    int64_t status = 0;

    /* vec![vec![0; m+1]; n+1] of usize (:175): n+1 separately allocated rows */
    size_t **table = (size_t **)malloc((n + 1) * sizeof(size_t *));
    for (size_t i = 0; i <= n; i++) {
        table[i] = (size_t *)malloc((m + 1) * sizeof(size_t));
        memset(table[i], 0, (m + 1) * sizeof(size_t));
    }

    /* forward fill (:176-187) */
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < m; j++) {
            uint32_t s = src_hash[i], d = dst_hash[j];
            if (i == 0 || j == 0) {
                table[i][j] = 0;
            } else if ((uint32_t)__builtin_popcount(s ^ d) <= threshold) {
                table[i][j] = table[i - 1][j - 1] + 1;
            } else {
                table[i][j] = 0;
            }
        }
    }

    /* reverse walk (:191-246) */
    for (size_t i = n - 1; i >= 1; i--) {
        for (size_t j = m - 1; j >= 1; j--) {
            if (table[i][j] == 0 || (i < n - 1 && j < m - 1 && table[i + 1][j + 1] != 0)) continue;

            size_t len = table[i][j];
            size_t src_start_idx = i - len, src_end_idx = i;
            size_t dst_start_idx = j - len, dst_end_idx = j;
            uint64_t src_start = src_ts[src_start_idx], src_end = src_ts[src_end_idx];
            uint64_t dst_start = dst_ts[dst_start_idx], dst_end = dst_ts[dst_end_idx];

            /* Duration - Duration panics on underflow */
            if (src_end < src_start || dst_end < dst_start) {
                status = -1;
                goto done;
            }
            uint64_t min_ns = is_opening ? min_opening_ns : min_ending_ns;
            int is_src_valid = (src_end - src_start) >= min_ns;
            int is_dst_valid = (dst_end - dst_start) >= min_ns;
            if (!(is_src_valid && is_dst_valid)) continue;

            orc_entry e;
            memset(&e, 0, sizeof e);
            e.score = len;
            e.src_start_ns = src_start;
            e.src_end_ns = src_end;
            e.dst_start_ns = dst_start;
            e.dst_end_ns = dst_end;
            /* compute_hash_for_match: simhash over hashes[start..end+1] (:149-153) */
            e.src_match_hash = orc_simhash32(src_hash + src_start_idx, src_end_idx - src_start_idx + 1);
            e.dst_match_hash = orc_simhash32(dst_hash + dst_start_idx, dst_end_idx - dst_start_idx + 1);
            e.is_ending = (uint32_t)is_ending;
            e.src_hash_duration_ns = src_hash_duration_ns;
            e.dst_hash_duration_ns = dst_hash_duration_ns;
            e.i_end = (uint32_t)i;
            e.j_end = (uint32_t)j;
            heap_push(&heap, &e);
        }
    }

done:
    for (size_t i = 0; i <= n; i++) free(table[i]);
    free(table);
    if (status < 0) {
        free(heap.data);
        return status;
    }
    *out = heap.data; /* heap.into(): the internal array, not sorted */
    return (int64_t)heap.len;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------- run_with_frame_hashes */

typedef struct {
    /* OpeningAndEndingInfo: src_openings == dst_openings and src_endings ==
     * dst_endings as entry lists (comparator.rs:283-300 pushes a clone to both
     * because longest_common_hash_match sets src and dst flags alike). */
    orc_entry *openings;
    int64_t n_openings;
    orc_entry *endings;
    int64_t n_endings;
    int status;
} pair_info;

typedef struct {
    const orc_season *season;
    const orc_params *params;
    const uint32_t (*pairs)[2];
    size_t n_pairs;
    pair_info *infos;
    size_t next; /* atomic work counter */
} search_job;

static void seg(const orc_season *s, size_t video, int ending, const uint32_t **h,
                const uint64_t **t, size_t *n) {
    uint64_t a = s->seg_offset[2 * video + (ending ? 1 : 0)];
    uint64_t b = s->seg_offset[2 * video + (ending ? 1 : 0) + 1];
    *h = s->hashes + a;
    *t = s->ts_ns + a;
    *n = (size_t)(b - a);
}

/* Comparator::search -> find_opening_and_ending (:383-399, :252-308) */
static void search_pair(const orc_season *s, const orc_params *p, size_t src, size_t dst,
                        pair_info *info) {
    const uint32_t *sh, *dh;
    const uint64_t *st, *dt;
    size_t n, m;
    memset(info, 0, sizeof *info);
    seg(s, src, 0, &sh, &st, &n);
    seg(s, dst, 0, &dh, &dt, &m);
    info->n_openings = orc_longest_common_hash_match(
        sh, st, n, dh, dt, m, p->hash_match_threshold, p->min_opening_ns, p->min_ending_ns,
        s->hash_duration_ns[src], s->hash_duration_ns[dst], 1, &info->openings);
    if (info->n_openings < 0) {
        info->status = ORC_ERR_UNDERFLOW;
        info->n_openings = 0;
        return;
    }
    if (p->include_endings) {
        seg(s, src, 1, &sh, &st, &n);
        seg(s, dst, 1, &dh, &dt, &m);
        if (n == 0 || m == 0) {
            info->status = ORC_ERR_NO_ENDING; /* :271-273 */
            return;
        }
        info->n_endings = orc_longest_common_hash_match(
            sh, st, n, dh, dt, m, p->hash_match_threshold, p->min_opening_ns, p->min_ending_ns,
            s->hash_duration_ns[src], s->hash_duration_ns[dst], 0, &info->endings);
        if (info->n_endings < 0) {
            info->status = ORC_ERR_UNDERFLOW;
            info->n_endings = 0;
        }
    }
}

static void *search_worker(void *arg) {
    search_job *job = (search_job *)arg;
    for (;;) {
        size_t k = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (k >= job->n_pairs) break;
        search_pair(job->season, job->params, job->pairs[k][0], job->pairs[k][1], &job->infos[k]);
    }
    return NULL;
}

typedef struct {
    uint64_t start_ns, end_ns, hash_duration_ns;
    uint32_t match_hash;
    int is_opening;
} candidate;

typedef struct {
    float score;
    size_t k;
} scored;

static int scored_cmp(const void *pa, const void *pb) {
    /* (f32, usize) partial_cmp, unwrap_or(Equal) (:473) */
    const scored *a = (const scored *)pa, *b = (const scored *)pb;
    if (a->score < b->score) return -1;
    if (a->score > b->score) return 1;
    if (a->score == b->score) {
        if (a->k < b->k) return -1;
        if (a->k > b->k) return 1;
    }
    return 0;
}

/* Comparator::find_best_match (:405-515).  matches = list of (info, is_source)
 * given as parallel arrays.  Returns ORC_OK / ORC_ERR_UNDERFLOW. */
static int find_best_match(const orc_params *p, pair_info *const *infos, const int *is_source,
                           size_t n_matches, orc_result *res) {
    memset(res, 0, sizeof *res);
    if (n_matches == 0) return ORC_OK; /* None */

    size_t n_cand = 0;
    for (size_t a = 0; a < n_matches; a++)
        n_cand += (size_t)(infos[a]->n_openings + infos[a]->n_endings);
    candidate *cand = (candidate *)malloc((n_cand ? n_cand : 1) * sizeof(candidate));
    size_t c = 0;
    for (size_t a = 0; a < n_matches; a++) {
        const pair_info *m = infos[a];
        for (int pass = 0; pass < 2; pass++) { /* openings, then endings (:413-431) */
            const orc_entry *list = pass == 0 ? m->openings : m->endings;
            int64_t cnt = pass == 0 ? m->n_openings : m->n_endings;
            for (int64_t e = 0; e < cnt; e++) {
                candidate *o = &cand[c++];
                if (is_source[a]) {
                    o->start_ns = list[e].src_start_ns;
                    o->end_ns = list[e].src_end_ns;
                    o->hash_duration_ns = list[e].src_hash_duration_ns;
                    o->match_hash = list[e].src_match_hash;
                } else {
                    o->start_ns = list[e].dst_start_ns;
                    o->end_ns = list[e].dst_end_ns;
                    o->hash_duration_ns = list[e].dst_hash_duration_ns;
                    o->match_hash = list[e].dst_match_hash;
                }
                o->is_opening = pass == 0;
            }
        }
    }

    /* distinct_matches: HashMap<usize, HashSet<usize>> (:434-454); only the
     * set sizes and key membership are used. */
    uint32_t bias = p->hash_match_threshold + p->hash_match_threshold / 2;
    size_t *count = (size_t *)calloc(n_cand ? n_cand : 1, sizeof(size_t));
    for (size_t i = 0; i < n_cand; i++) {
        for (size_t j = 0; j < n_cand; j++) {
            uint32_t dist = (uint32_t)__builtin_popcount(cand[i].match_hash ^ cand[j].match_hash);
            if (dist >= bias) continue;
            /* entry(i).insert(j); entry(j).insert(i): dist is symmetric, so
             * set(i) = { j : dist(i,j) < bias } and each (i,j) is inserted once
             * into set(i) as a distinct element. */
            count[i]++;
        }
    }

    res->present = 1;
    int status = ORC_OK;
    for (int want_opening = 1; want_opening >= 0; want_opening--) {
        if (!want_opening && !p->include_endings) break; /* :486 */
        scored *best = (scored *)malloc((n_cand ? n_cand : 1) * sizeof(scored));
        size_t nb = 0;
        for (size_t k = 0; k < n_cand; k++) {
            if (count[k] == 0) continue; /* not a key of distinct_matches */
            if (cand[k].is_opening != want_opening) continue;
            if (cand[k].end_ns < cand[k].start_ns) {
                status = ORC_ERR_UNDERFLOW;
                continue;
            }
            volatile float cnt_f = (float)(int64_t)count[k];
            volatile float dur = orc_duration_as_secs_f32(cand[k].end_ns - cand[k].start_ns);
            volatile float a = cnt_f * 0.3f;
            volatile float b = dur * 0.7f;
            volatile float s = a + b;
            best[nb].score = -s;
            best[nb].k = k;
            nb++;
        }
        qsort(best, nb, sizeof(scored), scored_cmp);
        if (nb > 0) {
            const candidate *w = &cand[best[0].k];
            uint64_t start = w->start_ns + p->time_padding_ns;
            uint64_t sub = p->time_padding_ns + w->hash_duration_ns;
            if (w->end_ns < sub) {
                status = ORC_ERR_UNDERFLOW; /* end - padding - hash_duration panics */
            } else if (want_opening) {
                res->has_opening = 1;
                res->opening_start_ns = start;
                res->opening_end_ns = w->end_ns - sub;
            } else {
                res->has_ending = 1;
                res->ending_start_ns = start;
                res->ending_end_ns = w->end_ns - sub;
            }
        }
        free(best);
    }
    free(count);
    free(cand);
    return status;
}

int orc_run_with_frame_hashes(const orc_season *season, const orc_params *params, int n_threads,
                              orc_result *results, orc_entry **entries_out,
                              uint32_t **entry_pair_out, uint64_t *n_entries_out) {
    size_t N = season->n_videos;
    /* pair list (:534-545): i ascending, j over not-yet-processed videos != i */
    size_t n_pairs = N * (N - (N ? 1 : 0)) / 2;
    uint32_t(*pairs)[2] = (uint32_t(*)[2])malloc((n_pairs ? n_pairs : 1) * sizeof *pairs);
    {
        char *processed = (char *)calloc(N ? N : 1, 1);
        size_t k = 0;
        for (size_t i = 0; i < N; i++) {
            for (size_t j = 0; j < N; j++) {
                if (i == j || processed[j]) continue;
                pairs[k][0] = (uint32_t)i;
                pairs[k][1] = (uint32_t)j;
                k++;
            }
            processed[i] = 1;
        }
        free(processed);
    }

    pair_info *infos = (pair_info *)calloc(n_pairs ? n_pairs : 1, sizeof(pair_info));
    search_job job = {season, params, (const uint32_t(*)[2])pairs, n_pairs, infos, 0};
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1) {
        search_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, search_worker, &job);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        free(th);
    }

    int status = ORC_OK;
    for (size_t k = 0; k < n_pairs; k++)
        if (infos[k].status != ORC_OK && status == ORC_OK) status = infos[k].status;

    /* info_map (:580-588): non-empty infos only, in pair order */
    size_t *n_match = (size_t *)calloc(N ? N : 1, sizeof(size_t));
    for (size_t k = 0; k < n_pairs; k++) {
        if (infos[k].n_openings + infos[k].n_endings == 0) continue;
        n_match[pairs[k][0]]++;
        n_match[pairs[k][1]]++;
    }
    if (status == ORC_OK) {
        for (size_t v = 0; v < N; v++) {
            pair_info **minfo = (pair_info **)malloc((n_match[v] ? n_match[v] : 1) * sizeof *minfo);
            int *msrc = (int *)malloc((n_match[v] ? n_match[v] : 1) * sizeof(int));
            size_t c = 0;
            for (size_t k = 0; k < n_pairs; k++) {
                if (infos[k].n_openings + infos[k].n_endings == 0) continue;
                if (pairs[k][0] == v) {
                    minfo[c] = &infos[k];
                    msrc[c++] = 1;
                } else if (pairs[k][1] == v) {
                    minfo[c] = &infos[k];
                    msrc[c++] = 0;
                }
            }
            int st = find_best_match(params, minfo, msrc, c, &results[v]);
            if (st != ORC_OK && status == ORC_OK) status = st;
            free(minfo);
            free(msrc);
        }
    }
    free(n_match);

    if (entries_out) {
        uint64_t total = 0;
        for (size_t k = 0; k < n_pairs; k++) total += (uint64_t)(infos[k].n_openings + infos[k].n_endings);
        orc_entry *all = (orc_entry *)malloc((total ? total : 1) * sizeof(orc_entry));
        uint32_t *pidx = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
        uint64_t c = 0;
        for (size_t k = 0; k < n_pairs; k++) {
            for (int64_t e = 0; e < infos[k].n_openings; e++) {
                all[c] = infos[k].openings[e];
                pidx[c++] = (uint32_t)k;
            }
            for (int64_t e = 0; e < infos[k].n_endings; e++) {
                all[c] = infos[k].endings[e];
                pidx[c++] = (uint32_t)k;
            }
        }
        *entries_out = all;
        if (entry_pair_out) *entry_pair_out = pidx; else free(pidx);
        if (n_entries_out) *n_entries_out = total;
    }

    for (size_t k = 0; k < n_pairs; k++) {
        free(infos[k].openings);
        free(infos[k].endings);
    }
    free(infos);
    free(pairs);
    return status;
}
