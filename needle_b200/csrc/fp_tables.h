// fp_tables.h -- constants of the Chromaprint TEST2 fingerprint algorithm
// (CHROMAPRINT_ALGORITHM_DEFAULT, what needle's chromaprint::Context::default()
// selects, needle/src/audio/analyzer.rs:176).  The arithmetic lives in the
// un-vendored crate chromaprint-sys-next 1.5.3 (needle/Cargo.lock:158-159):
// these are Chromaprint 1.5.x's fingerprinter_configuration.cpp values.
// Product copy; tests/test_tables.py checks it against the oracle's.
#pragma once

namespace nb200 {

constexpr int FP_SAMPLE_RATE = 11025;
constexpr int FP_FRAME = 4096;                       // kDefaultFrameSize
constexpr int FP_OVERLAP = FP_FRAME - FP_FRAME / 3;  // kDefaultFrameOverlap = 2731
constexpr int FP_HOP = FP_FRAME - FP_OVERLAP;        // 1365
constexpr int FP_BANDS = 12;
constexpr int FP_MIN_FREQ = 28;
constexpr int FP_MAX_FREQ = 3520;
constexpr int FP_FIR_LEN = 5;                        // kChromaFilterSize
constexpr int FP_WINDOW_ROWS = 16;                   // max filter width of the TEST2 classifiers
constexpr int FP_NUM_CLASSIFIERS = 16;
constexpr int FP_WARMUP = (FP_FIR_LEN - 1) + (FP_WINDOW_ROWS - 1);   // frames before the first hash
constexpr double FP_NORM_THRESHOLD = 0.01;

static const double FP_FIR_COEFFS[FP_FIR_LEN] = {0.25, 0.75, 1.0, 0.75, 0.25};

struct FpClassifierDef {
    int type, y, height, width;   // Filter(type, y, height, width)
    double t0, t1, t2;            // Quantizer(t0, t1, t2)
};

static const FpClassifierDef FP_CLASSIFIERS_TEST2[FP_NUM_CLASSIFIERS] = {
    {0, 4, 3, 15, 1.98215, 2.35817, 2.63523},
    {4, 4, 6, 15, -1.03809, -0.651211, -0.282167},
    {1, 0, 4, 16, -0.298702, 0.119262, 0.558497},
    {3, 8, 2, 12, -0.105439, 0.0153946, 0.135898},
    {3, 4, 4, 8, -0.142891, 0.0258736, 0.200632},
    {4, 0, 3, 5, -0.826319, -0.590612, -0.368214},
    {1, 2, 2, 9, -0.557409, -0.233035, 0.0534525},
    {2, 7, 3, 4, -0.0646826, 0.00620476, 0.0784847},
    {2, 6, 2, 16, -0.192387, -0.029699, 0.215855},
    {2, 1, 3, 2, -0.0397818, -0.00568076, 0.0292026},
    {5, 10, 1, 15, -0.53823, -0.369934, -0.190235},
    {3, 6, 2, 10, -0.124877, 0.0296483, 0.139239},
    {2, 1, 1, 14, -0.101475, 0.0225617, 0.231971},
    {3, 5, 6, 4, -0.0799915, -0.00729616, 0.063262},
    {1, 9, 2, 12, -0.272556, 0.019424, 0.302559},
    {3, 4, 2, 14, -0.164292, -0.0321188, 0.0846339},
};

}  // namespace nb200
