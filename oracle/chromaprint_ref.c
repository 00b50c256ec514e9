/*
 * chromaprint_ref.c -- double-precision CPU restatement of the Chromaprint
 * TEST2 fingerprinter that needle's Analyzer::process_frames drives through
 * chromaprint_rust::Context (needle/src/audio/analyzer.rs:176,218,275,286-301).
 * TEST INFRASTRUCTURE (see needle_oracle.h).  PARITY UNPINNED: Chromaprint's
 * sources are an un-vendored dependency (chromaprint-sys-next 1.5.3,
 * needle/Cargo.lock:158-159); this follows its published algorithm
 * [UPSTREAM-RECALL], file by file:
 *
 *   audio_processor.cpp   stereo -> mono (L+R)/2 in int arithmetic, no resample at 11025 Hz
 *   fft.cpp/audio_slicer  frames of 4096 samples, hop 1365, no tail padding
 *   fft_lib_*.cpp         Hamming window scaled by 1/INT16_MAX, |X[k]|^2 for k = 0..2048
 *   chroma.cpp            bins [10,1308) folded into 12 pitch classes, no interpolation
 *   chroma_filter.cpp     5-tap FIR {.25,.75,1,.75,.25} along time
 *   chroma_normalizer.h   Euclidean norm, zeroed if < 0.01
 *   fingerprint_calculator.cpp + utils/rolling_integral_image.h + filter.h +
 *   filter_utils.h + quantizer.h + utils/gray_code.h   16 classifiers -> u32
 */
#include "needle_oracle.h"
#include "chromaprint_tables.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

size_t orc_num_frames(size_t n) {
    return n >= ORC_FRAME_SIZE ? (n - ORC_FRAME_SIZE) / ORC_FRAME_HOP + 1 : 0;
}

size_t orc_num_raw_hashes(size_t n) {
    size_t f = orc_num_frames(n);
    /* 4 frames of FIR warm-up + 15 rows before the first 16-row window */
    size_t warm = (ORC_CHROMA_FILTER_LEN - 1) + (ORC_MAX_FILTER_WIDTH - 1);
    return f > warm ? f - warm : 0;
}

/* ------------------------------------------------------------------- FFT */

typedef struct {
    double re, im;
} cplx;

#define HALF (ORC_FRAME_SIZE / 2)

static double g_window[ORC_FRAME_SIZE];
static cplx g_tw_half[HALF];            /* exp(-2 pi i k / 2048) */
static cplx g_tw_full[HALF + 1];        /* exp(-2 pi i k / 4096), k = 0..2048 */
static signed char g_notes[ORC_FRAME_SIZE];
static int g_min_index, g_max_index;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables(void) {
    /* PrepareHammingWindow(first, last, 1.0 / INT16_MAX) */
    for (int i = 0; i < ORC_FRAME_SIZE; i++)
        g_window[i] = (1.0 / 32767.0) * (0.54 - 0.46 * cos(i * 2.0 * M_PI / (ORC_FRAME_SIZE - 1)));
    for (int k = 0; k < HALF; k++) {
        g_tw_half[k].re = cos(2.0 * M_PI * k / HALF);
        g_tw_half[k].im = -sin(2.0 * M_PI * k / HALF);
    }
    for (int k = 0; k <= HALF; k++) {
        g_tw_full[k].re = cos(2.0 * M_PI * k / ORC_FRAME_SIZE);
        g_tw_full[k].im = -sin(2.0 * M_PI * k / ORC_FRAME_SIZE);
    }
    orc_chroma_notes(ORC_MIN_FREQ, ORC_MAX_FREQ, ORC_FRAME_SIZE, ORC_SAMPLE_RATE, &g_min_index,
                     &g_max_index, g_notes);
}

/* Stockham autosort complex FFT of length 2048 = 4^5 * 2 (forward, e^{-i..}).
 * x holds the input and receives the output; y is scratch. */
static void fft2048(cplx *x, cplx *y) {
    int n = HALF, s = 1;
    cplx *in = x, *out = y;
    while (n >= 4) {
        int n1 = n / 4, n2 = n / 2, n3 = n1 + n2;
        for (int p = 0; p < n1; p++) {
            cplx w1 = g_tw_half[(p * s) & (HALF - 1)];
            cplx w2 = g_tw_half[(2 * p * s) & (HALF - 1)];
            cplx w3 = g_tw_half[(3 * p * s) & (HALF - 1)];
            const cplx *a_ = in + s * p, *b_ = in + s * (p + n1), *c_ = in + s * (p + n2),
                       *d_ = in + s * (p + n3);
            cplx *o0 = out + s * (4 * p), *o1 = o0 + s, *o2 = o1 + s, *o3 = o2 + s;
            for (int q = 0; q < s; q++) {
                cplx a = a_[q], b = b_[q], c = c_[q], d = d_[q];
                double apc_r = a.re + c.re, apc_i = a.im + c.im;
                double amc_r = a.re - c.re, amc_i = a.im - c.im;
                double bpd_r = b.re + d.re, bpd_i = b.im + d.im;
                /* j*(b - d) with forward transform sign: multiply by -i -> (im, -re) */
                double jr = (b.im - d.im), ji = -(b.re - d.re);
                o0[q].re = apc_r + bpd_r;
                o0[q].im = apc_i + bpd_i;
                double t1r = amc_r + jr, t1i = amc_i + ji;
                double t2r = apc_r - bpd_r, t2i = apc_i - bpd_i;
                double t3r = amc_r - jr, t3i = amc_i - ji;
                o1[q].re = t1r * w1.re - t1i * w1.im;
                o1[q].im = t1r * w1.im + t1i * w1.re;
                o2[q].re = t2r * w2.re - t2i * w2.im;
                o2[q].im = t2r * w2.im + t2i * w2.re;
                o3[q].re = t3r * w3.re - t3i * w3.im;
                o3[q].im = t3r * w3.im + t3i * w3.re;
            }
        }
        n /= 4;
        s *= 4;
        cplx *t = in;
        in = out;
        out = t;
    }
    if (n == 2) {
        for (int q = 0; q < s; q++) {
            cplx a = in[q], b = in[q + s];
            out[q].re = a.re + b.re;
            out[q].im = a.im + b.im;
            out[q + s].re = a.re - b.re;
            out[q + s].im = a.im - b.im;
        }
        cplx *t = in;
        in = out;
        out = t;
    }
    if (in != x) memcpy(x, in, HALF * sizeof(cplx));
}

/* power[k] = |X[k]|^2, k = 0..2048, X = unnormalised real DFT of the windowed
 * frame (FFTW r2hc / av_rdft convention). */
static void power_spectrum_mono(const int16_t *frame, double *power, cplx *z, cplx *scratch) {
    for (int n = 0; n < HALF; n++) {
        z[n].re = (double)frame[2 * n] * g_window[2 * n];
        z[n].im = (double)frame[2 * n + 1] * g_window[2 * n + 1];
    }
    fft2048(z, scratch);
    for (int k = 0; k <= HALF; k++) {
        cplx a = z[k & (HALF - 1)];
        cplx b = z[(HALF - k) & (HALF - 1)];
        /* E = (a + conj b)/2, O = (a - conj b)/(2i) */
        double er = 0.5 * (a.re + b.re), ei = 0.5 * (a.im - b.im);
        double or_ = 0.5 * (a.im + b.im), oi = -0.5 * (a.re - b.re);
        cplx w = g_tw_full[k];
        double xr = er + (or_ * w.re - oi * w.im);
        double xi = ei + (or_ * w.im + oi * w.re);
        power[k] = xr * xr + xi * xi;
    }
}

void orc_power_spectrum(const int16_t *frame, double *power) {
    pthread_once(&g_once, init_tables);
    cplx *z = (cplx *)malloc(2 * HALF * sizeof(cplx));
    power_spectrum_mono(frame, power, z, z + HALF);
    free(z);
}

/* ---------------------------------------------------------------- chroma */

static int freq_to_index(int freq, int frame_size, int sample_rate) {
    return (int)round((double)frame_size * freq / sample_rate);
}

void orc_chroma_notes(int min_freq, int max_freq, int frame_size, int sample_rate, int *min_index,
                      int *max_index, signed char *notes) {
    int lo = freq_to_index(min_freq, frame_size, sample_rate);
    int hi = freq_to_index(max_freq, frame_size, sample_rate);
    if (lo < 1) lo = 1;
    if (hi > frame_size / 2) hi = frame_size / 2;
    memset(notes, 0, (size_t)frame_size);
    for (int i = lo; i < hi; i++) {
        double freq = ((double)i * sample_rate) / frame_size;
        double octave = log(freq / (440.0 / 16.0)) / log(2.0);
        double note = ORC_NUM_BANDS * (octave - floor(octave));
        notes[i] = (signed char)note;
    }
    *min_index = lo;
    *max_index = hi;
}

void orc_chroma_fold(const double *power, int frame_size, int min_freq, int max_freq,
                     int sample_rate, double *features) {
    signed char *notes = (signed char *)malloc((size_t)frame_size);
    int lo, hi;
    orc_chroma_notes(min_freq, max_freq, frame_size, sample_rate, &lo, &hi, notes);
    for (int b = 0; b < ORC_NUM_BANDS; b++) features[b] = 0.0;
    for (int i = lo; i < hi; i++) features[notes[i]] += power[i];
    free(notes);
}

size_t orc_chroma_filter(const double *coeffs, int len, const double *rows, size_t n_rows,
                         double *out) {
    if (n_rows < (size_t)len) return 0;
    size_t n_out = n_rows - (size_t)len + 1;
    for (size_t t = 0; t < n_out; t++) {
        for (int b = 0; b < ORC_NUM_BANDS; b++) {
            double acc = 0.0;
            for (int j = 0; j < len; j++) acc += rows[(t + (size_t)j) * ORC_NUM_BANDS + b] * coeffs[j];
            out[t * ORC_NUM_BANDS + b] = acc;
        }
    }
    return n_out;
}

void orc_normalize(double *f, double threshold) {
    double squares = 0.0;
    for (int b = 0; b < ORC_NUM_BANDS; b++) squares += f[b] * f[b];
    double norm = squares > 0 ? sqrt(squares) : 0.0;
    if (norm < threshold) {
        for (int b = 0; b < ORC_NUM_BANDS; b++) f[b] = 0.0;
    } else {
        for (int b = 0; b < ORC_NUM_BANDS; b++) f[b] /= norm;
    }
}

/* ------------------------------------------------- classifiers / image */

int orc_quantize(double v, double t0, double t1, double t2) {
    if (v < t1) return v < t0 ? 0 : 1;
    return v < t2 ? 2 : 3;
}

int orc_gray_code(int i) {
    static const int codes[4] = {0, 1, 3, 2};
    return codes[i];
}

/* Integral image over a full n_rows x n_cols array, built the way
 * RollingIntegralImage::AddRow does: partial_sum along the row, then add the
 * previous integral row. */
static void integral_image(const double *image, size_t n_rows, size_t n_cols, double *integ) {
    for (size_t r = 0; r < n_rows; r++) {
        double acc = 0.0;
        for (size_t c = 0; c < n_cols; c++) {
            acc += image[r * n_cols + c];
            integ[r * n_cols + c] = acc;
        }
        if (r > 0)
            for (size_t c = 0; c < n_cols; c++) integ[r * n_cols + c] += integ[(r - 1) * n_cols + c];
    }
}

/* RollingIntegralImage::Area(r1, c1, r2, c2): rows [r1,r2) x cols [c1,c2) */
static double area(const double *integ, size_t n_cols, size_t r1, size_t c1, size_t r2, size_t c2) {
    if (r1 == r2 || c1 == c2) return 0.0;
    if (r1 == 0) {
        const double *row = integ + (r2 - 1) * n_cols;
        return c1 == 0 ? row[c2 - 1] : row[c2 - 1] - row[c1 - 1];
    }
    const double *row1 = integ + (r1 - 1) * n_cols;
    const double *row2 = integ + (r2 - 1) * n_cols;
    if (c1 == 0) return row2[c2 - 1] - row1[c2 - 1];
    return row2[c2 - 1] - row1[c2 - 1] - row2[c1 - 1] + row1[c1 - 1];
}

static double subtract_log(double a, double b) { return log((1.0 + a) / (1.0 + b)); }

static double filter_on_integral(int type, size_t y, size_t h, size_t w, const double *integ,
                                 size_t n_cols, size_t x) {
    double a = 0, b = 0;
    switch (type) {
    case 0:
        a = area(integ, n_cols, x, y, x + w, y + h);
        b = 0;
        break;
    case 1: {
        size_t h2 = h / 2;
        a = area(integ, n_cols, x, y + h2, x + w, y + h);
        b = area(integ, n_cols, x, y, x + w, y + h2);
        break;
    }
    case 2: {
        size_t w2 = w / 2;
        a = area(integ, n_cols, x + w2, y, x + w, y + h);
        b = area(integ, n_cols, x, y, x + w2, y + h);
        break;
    }
    case 3: {
        size_t w2 = w / 2, h2 = h / 2;
        a = area(integ, n_cols, x, y + h2, x + w2, y + h) +
            area(integ, n_cols, x + w2, y, x + w, y + h2);
        b = area(integ, n_cols, x, y, x + w2, y + h2) +
            area(integ, n_cols, x + w2, y + h2, x + w, y + h);
        break;
    }
    case 4: {
        size_t h3 = h / 3;
        a = area(integ, n_cols, x, y + h3, x + w, y + 2 * h3);
        b = area(integ, n_cols, x, y, x + w, y + h3) +
            area(integ, n_cols, x, y + 2 * h3, x + w, y + h);
        break;
    }
    case 5: {
        size_t w3 = w / 3;
        a = area(integ, n_cols, x + w3, y, x + 2 * w3, y + h);
        b = area(integ, n_cols, x, y, x + w3, y + h) +
            area(integ, n_cols, x + 2 * w3, y, x + w, y + h);
        break;
    }
    }
    return subtract_log(a, b);
}

double orc_filter_apply(int type, int y, int height, int width, const double *image,
                        size_t n_rows, size_t n_cols, size_t x) {
    double *integ = (double *)malloc(n_rows * n_cols * sizeof(double));
    integral_image(image, n_rows, n_cols, integ);
    double v = filter_on_integral(type, (size_t)y, (size_t)height, (size_t)width, integ, n_cols, x);
    free(integ);
    return v;
}

/* ------------------------------------------------------------ fingerprint */

int64_t orc_fingerprint(const int16_t *pcm, size_t n_samples_total, int channels, uint32_t **out,
                        double **chroma_out) {
    *out = NULL;
    if (chroma_out) *chroma_out = NULL;
    if (channels != 1 && channels != 2) return -1;
    if (n_samples_total % (size_t)channels) return -1;
    pthread_once(&g_once, init_tables);

    size_t n = n_samples_total / (size_t)channels;
    size_t n_frames = orc_num_frames(n);
    size_t n_raw = orc_num_raw_hashes(n);

    /* AudioProcessor::LoadMono / LoadStereo */
    int16_t *mono = (int16_t *)malloc((n ? n : 1) * sizeof(int16_t));
    if (channels == 1) {
        memcpy(mono, pcm, n * sizeof(int16_t));
    } else {
        for (size_t i = 0; i < n; i++) mono[i] = (int16_t)(((int)pcm[2 * i] + (int)pcm[2 * i + 1]) / 2);
    }

    double *chroma = (double *)malloc((n_frames ? n_frames : 1) * ORC_NUM_BANDS * sizeof(double));
    {
        double *power = (double *)malloc((HALF + 1) * sizeof(double));
        cplx *z = (cplx *)malloc(2 * HALF * sizeof(cplx));
        for (size_t f = 0; f < n_frames; f++) {
            power_spectrum_mono(mono + f * ORC_FRAME_HOP, power, z, z + HALF);
            double *feat = chroma + f * ORC_NUM_BANDS;
            for (int b = 0; b < ORC_NUM_BANDS; b++) feat[b] = 0.0;
            for (int i = g_min_index; i < g_max_index; i++) feat[g_notes[i]] += power[i];
        }
        free(power);
        free(z);
    }
    free(mono);

    uint32_t *hashes = (uint32_t *)malloc((n_raw ? n_raw : 1) * sizeof(uint32_t));
    if (n_raw > 0) {
        size_t n_rows = n_frames - (ORC_CHROMA_FILTER_LEN - 1);
        double *rows = (double *)malloc(n_rows * ORC_NUM_BANDS * sizeof(double));
        orc_chroma_filter(orc_chroma_filter_coeffs, ORC_CHROMA_FILTER_LEN, chroma, n_frames, rows);
        for (size_t r = 0; r < n_rows; r++) orc_normalize(rows + r * ORC_NUM_BANDS, 0.01);
        double *integ = (double *)malloc(n_rows * ORC_NUM_BANDS * sizeof(double));
        integral_image(rows, n_rows, ORC_NUM_BANDS, integ);
        /* FingerprintCalculator::Consume: once num_rows >= 16, one sub-fingerprint
         * per row at offset num_rows - 16 */
        for (size_t x = 0; x < n_raw; x++) {
            uint32_t bits = 0;
            for (int c = 0; c < ORC_NUM_CLASSIFIERS; c++) {
                const orc_classifier *k = &orc_classifiers_test2[c];
                double v = filter_on_integral(k->type, (size_t)k->y, (size_t)k->height,
                                              (size_t)k->width, integ, ORC_NUM_BANDS, x);
                bits = (bits << 2) | (uint32_t)orc_gray_code(orc_quantize(v, k->t0, k->t1, k->t2));
            }
            hashes[x] = bits;
        }
        free(integ);
        free(rows);
    }

    if (chroma_out) *chroma_out = chroma; else free(chroma);
    *out = hashes;
    return (int64_t)n_raw;
}

size_t orc_subsample_and_stamp(const uint32_t *raw, size_t n_raw, uint32_t step_by,
                               uint64_t delay_ns, uint64_t item_ns, uint64_t seek_to_ns,
                               uint32_t *out_hash, uint64_t *out_ts) {
    size_t c = 0;
    if (step_by == 0) return 0; /* Rust's step_by(0) panics */
    for (size_t i = 0; i < n_raw; i += step_by) {
        out_hash[c] = raw[i];
        out_ts[c] = orc_hash_timestamp(delay_ns, item_ns, (uint32_t)i, seek_to_ns);
        c++;
    }
    return c;
}

/* ------------------------------------------------------- threaded driver */

typedef struct {
    const int16_t *const *pcm;
    const uint64_t *n_samples_total;
    int channels;
    size_t n_segments;
    uint32_t **out;
    uint64_t *out_counts;
    size_t next;
    int status;
} fp_job;

static void *fp_worker(void *arg) {
    fp_job *job = (fp_job *)arg;
    for (;;) {
        size_t k = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (k >= job->n_segments) break;
        int64_t n = orc_fingerprint(job->pcm[k], (size_t)job->n_samples_total[k], job->channels,
                                    &job->out[k], NULL);
        if (n < 0) {
            job->status = -1;
            job->out_counts[k] = 0;
        } else {
            job->out_counts[k] = (uint64_t)n;
        }
    }
    return NULL;
}

int orc_fingerprint_many(const int16_t *const *pcm, const uint64_t *n_samples_total, int channels,
                         size_t n_segments, int n_threads, uint32_t **out, uint64_t *out_counts) {
    fp_job job = {pcm, n_samples_total, channels, n_segments, out, out_counts, 0, 0};
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1) {
        fp_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, fp_worker, &job);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    return job.status;
}
