//! needle/src/audio/gpu.rs -- safe wrappers over ffi.rs, shaped like what they replace.
//!
//! * `Fingerprinter` has the methods of `chromaprint_rust::Context` that
//!   `Analyzer::process_frames` calls (analyzer.rs:176,179,218,275,286,288,289,299-301), with
//!   the same names and return shapes, so the patch there is a type swap.
//! * `match_pairs` is the body of the pair loop of `Comparator::run_with_frame_hashes`
//!   (comparator.rs:547-578): one call for all pairs; comparator.rs rebuilds its private
//!   `ComparatorHeapEntry`s from the returned runs (integration/comparator.patch).
//! * `MultiGpu` drives every GPU of the box from this one process.
use std::cell::RefCell;
use std::ffi::CStr;
use std::time::Duration;

use super::ffi::*;
use super::FrameHashes;
use crate::{Error, Result};

fn check(status: i32) -> Result<()> {
    match status {
        NB200_OK => Ok(()),
        NB200_ERR_NO_ENDING => Err(Error::FrameHashDataNoEnding),
        NB200_ERR_FORMAT => Err(Error::FrameHashDataInvalidVersion),
        s => {
            // NB200_ERR_DURATION_UNDERFLOW is where the reference panics on `Duration - Duration`
            let what = unsafe { CStr::from_ptr(nb200_status_str(s)) }.to_string_lossy().into_owned();
            let detail = unsafe { CStr::from_ptr(nb200_last_error()) }.to_string_lossy().into_owned();
            Err(Error::Gpu(format!("{what} {detail}")))
        }
    }
}

/// One library context (CUDA stream + scratch) per rayon worker thread: contexts are
/// independent, calls on one context must not overlap.
pub struct Context(*mut nb200_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { nb200_ctx_create(device, &mut p) })?;
        Ok(Context(p))
    }
    pub fn raw(&self) -> *mut nb200_ctx {
        self.0
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { nb200_ctx_destroy(self.0) }
    }
}

thread_local! {
    static CTX: RefCell<Option<Context>> = RefCell::new(None);
}

/// The calling thread's context (device: NEEDLE_B200_DEVICE, default the current one).
pub fn with_context<T>(f: impl FnOnce(&Context) -> Result<T>) -> Result<T> {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let dev = std::env::var("NEEDLE_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(-1);
            *c = Some(Context::new(dev)?);
        }
        f(c.as_ref().unwrap())
    })
}

/// Borrowed view of the raw fingerprint, like chromaprint_rust's `RawFingerprint::get()`.
pub struct RawFingerprint<'a>(&'a [u32]);
impl<'a> RawFingerprint<'a> {
    pub fn get(&self) -> &'a [u32] {
        self.0
    }
}

/// Drop-in for `chromaprint_rust::Context` inside `process_frames`.
pub struct Fingerprinter(*mut nb200_fp);

impl Fingerprinter {
    /// `chromaprint::Context::default()` (analyzer.rs:176)
    pub fn new() -> Result<Self> {
        with_context(|ctx| {
            let mut p = std::ptr::null_mut();
            check(unsafe { nb200_fp_new(ctx.raw(), &mut p) })?;
            Ok(Fingerprinter(p))
        })
    }
    /// `.sample_rate()` (:179) -- 11025
    pub fn sample_rate(&self) -> u32 {
        unsafe { nb200_fp_sample_rate(self.0) as u32 }
    }
    /// `.start(rate, 2)` (:218)
    pub fn start(&mut self, sample_rate: u32, channels: u16) -> Result<()> {
        check(unsafe { nb200_fp_start(self.0, sample_rate as i32, channels as i32) })
    }
    /// `.feed(&[i16])` (:275): accumulated in pinned host memory
    pub fn feed(&mut self, samples: &[i16]) -> Result<()> {
        check(unsafe { nb200_fp_feed(self.0, samples.as_ptr(), samples.len()) })
    }
    /// `.finish()` (:286): H2D, down-mix, K1, K2, raw hashes back
    pub fn finish(&mut self) -> Result<()> {
        check(unsafe { nb200_fp_finish(self.0) })
    }
    /// `.get_delay()` (:288)
    pub fn get_delay(&self) -> Result<Duration> {
        let mut ms = 0;
        check(unsafe { nb200_fp_get_delay_ms(self.0, &mut ms) })?;
        Ok(Duration::from_millis(ms as u64))
    }
    /// `.get_item_duration()` (:289)
    pub fn get_item_duration(&self) -> Result<Duration> {
        let mut ms = 0;
        check(unsafe { nb200_fp_get_item_duration_ms(self.0, &mut ms) })?;
        Ok(Duration::from_millis(ms as u64))
    }
    /// `.get_fingerprint_raw()?.get()` (:299-301)
    pub fn get_fingerprint_raw(&self) -> Result<RawFingerprint<'_>> {
        let (mut p, mut n) = (std::ptr::null(), 0usize);
        check(unsafe { nb200_fp_get_raw(self.0, &mut p, &mut n) })?;
        Ok(RawFingerprint(if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(p, n) } }))
    }
}
impl Drop for Fingerprinter {
    fn drop(&mut self) {
        unsafe { nb200_fp_free(self.0) }
    }
}

/// A season in the SoA form the library takes (FrameHashes::opening_data / ending_data,
/// data.rs:143-155): segment 2k = video k's opening list, 2k+1 its ending list.
pub struct Season {
    pub hashes: Vec<u32>,
    pub ts_ns: Vec<u64>,
    pub seg_offset: Vec<u64>,
    pub hash_duration_ns: Vec<u64>,
}

// The library stages ordinary memory through its own pinned area (about 1 ms per 10 MB).  Arrays that
// are page-locked already (nb200_host_alloc / nb200_host_free: a 20-line `PinnedVec<T>` around them) are
// copied from directly -- 6.3 -> 5.0 ms for a 19,900-pair nb200_search.  Plain Vecs keep this file short.
impl Season {
    pub fn from_frame_hashes(frame_hashes: &[FrameHashes]) -> Self {
        let (mut hashes, mut ts_ns, mut seg_offset) = (Vec::new(), Vec::new(), vec![0u64]);
        for f in frame_hashes {
            for list in [f.opening_data(), f.ending_data()] {
                for (h, d) in list {
                    hashes.push(*h);
                    ts_ns.push(d.as_nanos() as u64);
                }
                seg_offset.push(hashes.len() as u64);
            }
        }
        let hash_duration_ns = frame_hashes.iter().map(|f| f.hash_duration().as_nanos() as u64).collect();
        Season { hashes, ts_ns, seg_offset, hash_duration_ns }
    }
    pub fn n_videos(&self) -> u32 {
        ((self.seg_offset.len() - 1) / 2) as u32
    }
}

/// Every pair's runs in the reference's push order (pair asc, opening before ending, i desc,
/// j desc: comparator.rs:191-192).  `pairs` must be the list run_with_frame_hashes built (:534-545).
pub fn match_pairs(season: &Season, pairs: &[(usize, usize)], params: &nb200_match_params) -> Result<Vec<nb200_run>> {
    let flat: Vec<[u32; 2]> = pairs.iter().map(|&(a, b)| [a as u32, b as u32]).collect();
    with_context(|ctx| {
        let (mut runs, mut n) = (std::ptr::null_mut(), 0u64);
        check(unsafe {
            nb200_match_pairs(ctx.raw(), season.hashes.as_ptr(), season.ts_ns.as_ptr(), season.seg_offset.as_ptr(),
                              season.n_videos(), flat.as_ptr(), flat.len() as u64, params, &mut runs, &mut n)
        })?;
        let out = if n == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(runs, n as usize) }.to_vec() };
        unsafe { nb200_free(runs as *mut _) };
        Ok(out)
    })
}

/// The whole of run_with_frame_hashes' compute (match, heap replay, find_best_match) on the
/// device: bit-identical intervals (needle-b200 tests/test_vote_gpu.py, test_full_size_gpu.py).
pub fn search(season: &Season, params: &nb200_match_params) -> Result<Vec<nb200_search_result>> {
    with_context(|ctx| {
        let mut res = vec![nb200_search_result::default(); season.n_videos() as usize];
        check(unsafe {
            nb200_search(ctx.raw(), season.hashes.as_ptr(), season.ts_ns.as_ptr(), season.seg_offset.as_ptr(),
                         season.hash_duration_ns.as_ptr(), season.n_videos(), params, res.as_mut_ptr())
        })?;
        Ok(res)
    })
}

/// Every GPU of the box from this one process (ncclCommInitAll inside the library): the pair
/// loop sharded over the devices, runs pushed to device 0 over NVLink, vote there.
pub struct MultiGpu {
    ctxs: Vec<Context>,
    comms: Vec<*mut nb200_comm>,
}

impl MultiGpu {
    pub fn new(devices: &[i32]) -> Result<Self> {
        let ctxs = devices.iter().map(|&d| Context::new(d)).collect::<Result<Vec<_>>>()?;
        let raw: Vec<*mut nb200_ctx> = ctxs.iter().map(|c| c.raw()).collect();
        let mut comms = vec![std::ptr::null_mut(); raw.len()];
        check(unsafe { nb200_comm_init_all(raw.as_ptr(), raw.len() as i32, comms.as_mut_ptr()) })?;
        Ok(MultiGpu { ctxs, comms })
    }

    pub fn search(&self, season: &Season, params: &nb200_match_params) -> Result<Vec<nb200_search_result>> {
        let mut job = std::ptr::null_mut();
        check(unsafe {
            nb200_mjob_search_create(self.comms.as_ptr(), self.comms.len() as i32, season.hashes.as_ptr(),
                                     season.ts_ns.as_ptr(), season.seg_offset.as_ptr(), season.n_videos(),
                                     season.hash_duration_ns.as_ptr(), std::ptr::null(), 0, params, &mut job)
        })?;
        let mut res = vec![nb200_search_result::default(); season.n_videos() as usize];
        let st = unsafe { nb200_mjob_run(job, std::ptr::null(), res.as_mut_ptr()) };
        unsafe { nb200_mjob_free(job) };
        check(st)?;
        Ok(res)
    }
}
impl Drop for MultiGpu {
    fn drop(&mut self) {
        for &c in &self.comms {
            unsafe { nb200_comm_destroy(c) }
        }
        let _ = &self.ctxs; // contexts outlive their comms
    }
}
