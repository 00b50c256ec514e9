//! needle/src/audio/ffi.rs -- raw declarations of include/needle_b200.h (the parts needle uses).
//!
//! Hand-written mirror of the C header; every function here is exercised through the same
//! C ABI by the Python test-suite of needle-b200 (tests/, ctypes), so signatures can be checked
//! against `needle_b200/_lib.py::PROTOTYPES` line by line.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const NB200_OK: c_int = 0;
pub const NB200_ERR_NULL_ARGUMENT: c_int = 1;
pub const NB200_ERR_INVALID_ARGUMENT: c_int = 2;
pub const NB200_ERR_CUDA: c_int = 3;
pub const NB200_ERR_NO_ENDING: c_int = 4;
pub const NB200_ERR_DURATION_UNDERFLOW: c_int = 5;
pub const NB200_ERR_TOO_LARGE: c_int = 6;
pub const NB200_ERR_IO: c_int = 7;
pub const NB200_ERR_FORMAT: c_int = 8;
pub const NB200_ERR_STATE: c_int = 9;
pub const NB200_ERR_COMPARATOR_MINIMUM_PATHS: c_int = 10;
pub const NB200_ERR_NCCL: c_int = 11;

pub const NB200_DELAY_MS: u64 = 2600; // chromaprint_get_delay_ms
pub const NB200_ITEM_DURATION_MS: u64 = 123; // chromaprint_get_item_duration_ms
pub const NB200_UNIQUE_ID_BYTES: usize = 128;

#[repr(C)]
pub struct nb200_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct nb200_fp {
    _p: [u8; 0],
}
#[repr(C)]
pub struct nb200_comm {
    _p: [u8; 0],
}
#[repr(C)]
pub struct nb200_mjob {
    _p: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct nb200_match_params {
    pub hash_match_threshold: u32,
    pub include_endings: u32,
    pub min_opening_ns: u64,
    pub min_ending_ns: u64,
    pub time_padding_ns: u64,
}

/// One ComparatorHeapEntry in index form (comparator.rs:20-35, built at :231-243).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct nb200_run {
    pub pair: u32,
    pub is_ending: u32,
    pub i_end: u32,
    pub j_end: u32,
    pub len: u32,
    pub src_simhash: u32,
    pub dst_simhash: u32,
    pub reserved: u32,
    pub src_start_ns: u64,
    pub src_end_ns: u64,
    pub dst_start_ns: u64,
    pub dst_end_ns: u64,
}

/// SearchResult (comparator.rs:65-69) with the video index kept.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct nb200_search_result {
    pub present: u32,
    pub has_opening: u32,
    pub has_ending: u32,
    pub reserved: u32,
    pub opening_start_ns: u64,
    pub opening_end_ns: u64,
    pub ending_start_ns: u64,
    pub ending_end_ns: u64,
}

extern "C" {
    pub fn nb200_status_str(status: c_int) -> *const c_char;
    pub fn nb200_last_error() -> *const c_char;
    pub fn nb200_ctx_create(device: c_int, out: *mut *mut nb200_ctx) -> c_int;
    pub fn nb200_ctx_destroy(ctx: *mut nb200_ctx);
    pub fn nb200_free(p: *mut c_void);

    // ---- B1: the chromaprint_rust::Context calls of Analyzer::process_frames (analyzer.rs:176-301)
    pub fn nb200_fp_new(ctx: *mut nb200_ctx, out: *mut *mut nb200_fp) -> c_int;
    pub fn nb200_fp_free(fp: *mut nb200_fp);
    pub fn nb200_fp_sample_rate(fp: *const nb200_fp) -> c_int;
    pub fn nb200_fp_start(fp: *mut nb200_fp, sample_rate: c_int, channels: c_int) -> c_int;
    pub fn nb200_fp_feed(fp: *mut nb200_fp, data: *const i16, n_samples_total: usize) -> c_int;
    pub fn nb200_fp_finish(fp: *mut nb200_fp) -> c_int;
    pub fn nb200_fp_get_delay_ms(fp: *const nb200_fp, out: *mut c_int) -> c_int;
    pub fn nb200_fp_get_item_duration_ms(fp: *const nb200_fp, out: *mut c_int) -> c_int;
    pub fn nb200_fp_get_raw(fp: *const nb200_fp, hashes: *mut *const u32, n: *mut usize) -> c_int;

    // ---- B2: the pair loop of Comparator::run_with_frame_hashes (comparator.rs:532-578)
    pub fn nb200_match_pairs(
        ctx: *mut nb200_ctx,
        hashes: *const u32,
        ts_ns: *const u64,
        seg_offset: *const u64,
        n_videos: u32,
        pairs: *const [u32; 2],
        n_pairs: u64,
        params: *const nb200_match_params,
        out_runs: *mut *mut nb200_run,
        out_n: *mut u64,
    ) -> c_int;
    pub fn nb200_search(
        ctx: *mut nb200_ctx,
        hashes: *const u32,
        ts_ns: *const u64,
        seg_offset: *const u64,
        hash_duration_ns: *const u64,
        n_videos: u32,
        params: *const nb200_match_params,
        results: *mut nb200_search_result,
    ) -> c_int;
    /// `needle search --analyze` in one call: PCM (2 segments per video) -> intervals.
    pub fn nb200_analyze_search(
        ctx: *mut nb200_ctx,
        pcm: *const *const i16,
        n_samples_total: *const u64,
        channels: c_int,
        n_videos: u32,
        seek_to_ns: *const u64,
        hash_duration_ns: u64,
        params: *const nb200_match_params,
        results: *mut nb200_search_result,
    ) -> c_int;

    // ---- all GPUs of the box from this one process (the rayon fan-outs, analyzer.rs:437-445,
    //      comparator.rs:549-564, across devices)
    pub fn nb200_comm_init_all(ctxs: *const *mut nb200_ctx, n: c_int, out: *mut *mut nb200_comm) -> c_int;
    pub fn nb200_comm_destroy(comm: *mut nb200_comm);
    pub fn nb200_mjob_search_create(
        comms: *const *mut nb200_comm,
        n_local: c_int,
        hashes: *const u32,
        ts_ns: *const u64,
        seg_offset: *const u64,
        n_videos: u32,
        hash_duration_ns: *const u64,
        pairs: *const [u32; 2],
        n_pairs: u64,
        params: *const nb200_match_params,
        out: *mut *mut nb200_mjob,
    ) -> c_int;
    pub fn nb200_mjob_season_create(
        comms: *const *mut nb200_comm,
        n_local: c_int,
        n_mono_samples: *const u64,
        seek_to_ns: *const u64,
        n_videos: u32,
        hash_duration_ns: u64,
        pairs: *const [u32; 2],
        n_pairs: u64,
        params: *const nb200_match_params,
        out: *mut *mut nb200_mjob,
    ) -> c_int;
    pub fn nb200_mjob_run(job: *mut nb200_mjob, host_pcm: *const *const i16, results: *mut nb200_search_result) -> c_int;
    pub fn nb200_mjob_free(job: *mut nb200_mjob);
}
