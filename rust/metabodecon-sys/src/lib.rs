//! Raw bindings to `include/mdb200.h` (ABI version 2).  One declaration per exported symbol;
//! see the header for the reference interface (file:line) each one replaces.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub type mdb_status = c_int;
pub const MDB_OK: mdb_status = 0;
pub const MDB_ERR_NO_PEAKS_DETECTED: mdb_status = 1;
pub const MDB_ERR_EMPTY_SIGNAL_REGION: mdb_status = 2;
pub const MDB_ERR_EMPTY_SIGNAL_FREE_REGION: mdb_status = 3;
pub const MDB_ERR_INVALID_SMOOTHING_SETTINGS: mdb_status = 4;
pub const MDB_ERR_INVALID_SELECTION_SETTINGS: mdb_status = 5;
pub const MDB_ERR_INVALID_FITTING_SETTINGS: mdb_status = 6;
pub const MDB_ERR_INVALID_IGNORE_REGION: mdb_status = 7;
pub const MDB_ERR_EMPTY_DATA: mdb_status = 10;
pub const MDB_ERR_DATA_LENGTH_MISMATCH: mdb_status = 11;
pub const MDB_ERR_NON_UNIFORM_SPACING: mdb_status = 12;
pub const MDB_ERR_INVALID_INTENSITIES: mdb_status = 13;
pub const MDB_ERR_INVALID_SIGNAL_BOUNDARIES: mdb_status = 14;
pub const MDB_ERR_REFERENCE_PANIC: mdb_status = 100;
pub const MDB_ERR_CUDA: mdb_status = 200;
pub const MDB_ERR_INVALID_ARGUMENT: mdb_status = 201;
pub const MDB_ERR_UNSUPPORTED: mdb_status = 202;

pub const MDB_SMOOTHING_IDENTITY: i32 = 0;
pub const MDB_SMOOTHING_MOVING_AVERAGE: i32 = 1;
pub const MDB_SELECTION_DETECTOR_ONLY: i32 = 0;
pub const MDB_SELECTION_NOISE_SCORE_FILTER: i32 = 1;
pub const MDB_SCORING_MINIMUM_SUM: i32 = 0;
pub const MDB_FITTING_ANALYTICAL: i32 = 0;
pub const MDB_MEM_HOST: c_int = 0;
pub const MDB_MEM_DEVICE: c_int = 1;
pub const MDB_SUPERPOSITION_EXACT: c_int = 0;
pub const MDB_SUPERPOSITION_FAST: c_int = 1;
pub const MDB_FIT_EXACT: c_int = 0;
pub const MDB_FIT_CORRECTED: c_int = 1;
pub const MDB_FIT_ULP: c_int = 2;

/// `Lorentzian {sfhw, hw2, maxp}` (metabodecon/src/deconvolution/lorentzian.rs:138-145).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct mdb_lorentzian {
    pub sfhw: f64,
    pub hw2: f64,
    pub maxp: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mdb_smoothing_settings {
    pub kind: i32,
    pub iterations: u64,
    pub window_size: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mdb_selection_settings {
    pub kind: i32,
    pub scoring_method: i32,
    pub threshold: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mdb_fitting_settings {
    pub kind: i32,
    pub iterations: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct mdb_spectrum_view {
    pub chemical_shifts: *const f64,
    pub intensities: *const f64,
    pub len: usize,
    pub signal_boundaries: [f64; 2],
}

#[repr(C)]
pub struct mdb_deconvoluter {
    _private: [u8; 0],
}

#[repr(C)]
pub struct mdb_batch {
    _private: [u8; 0],
}

extern "C" {
    pub fn mdb_abi_version() -> u32;
    pub fn mdb_last_error_message() -> *const c_char;
    pub fn mdb_device_count() -> c_int;
    pub fn mdb_set_device_count(n: c_int) -> mdb_status;
    pub fn mdb_set_superposition_mode(mode: c_int) -> mdb_status;
    pub fn mdb_superposition_mode() -> c_int;
    pub fn mdb_measure_fp64_rate(dfma_per_s: *mut f64, dadd_per_s: *mut f64) -> mdb_status;
    pub fn mdb_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> mdb_status;
    pub fn mdb_host_free(ptr: *mut c_void) -> mdb_status;
    pub fn mdb_release_workspaces() -> mdb_status;
    pub fn mdb_kernel_launch_count() -> u64;
    pub fn mdb_transfer_bytes(h2d: *mut u64, d2h: *mut u64);
    pub fn mdb_reset_kernel_launch_count();
    pub fn mdb_profile_enable(on: c_int);
    pub fn mdb_profile_reset();
    pub fn mdb_profile_read(kernel: c_int, ms: *mut f64, launches: *mut u64, work: *mut f64) -> mdb_status;

    pub fn mdb_spectrum_validate(
        chemical_shifts: *const f64,
        n_shifts: usize,
        intensities: *const f64,
        n_intensities: usize,
        signal_boundaries: *const f64,
        ordered_boundaries: *mut f64,
    ) -> mdb_status;

    pub fn mdb_deconvoluter_default(out: *mut *mut mdb_deconvoluter) -> mdb_status;
    pub fn mdb_deconvoluter_new(
        smoothing: *const mdb_smoothing_settings,
        selection: *const mdb_selection_settings,
        fitting: *const mdb_fitting_settings,
        out: *mut *mut mdb_deconvoluter,
    ) -> mdb_status;
    pub fn mdb_deconvoluter_clone(src: *const mdb_deconvoluter, out: *mut *mut mdb_deconvoluter) -> mdb_status;
    pub fn mdb_deconvoluter_free(d: *mut mdb_deconvoluter);
    pub fn mdb_deconvoluter_smoothing_settings(d: *const mdb_deconvoluter, out: *mut mdb_smoothing_settings) -> mdb_status;
    pub fn mdb_deconvoluter_selection_settings(d: *const mdb_deconvoluter, out: *mut mdb_selection_settings) -> mdb_status;
    pub fn mdb_deconvoluter_fitting_settings(d: *const mdb_deconvoluter, out: *mut mdb_fitting_settings) -> mdb_status;
    pub fn mdb_deconvoluter_ignore_regions(d: *const mdb_deconvoluter, pairs: *mut f64, cap: usize) -> i64;
    pub fn mdb_deconvoluter_set_smoothing_settings(d: *mut mdb_deconvoluter, s: *const mdb_smoothing_settings) -> mdb_status;
    pub fn mdb_deconvoluter_set_selection_settings(d: *mut mdb_deconvoluter, s: *const mdb_selection_settings) -> mdb_status;
    pub fn mdb_deconvoluter_set_fitting_settings(d: *mut mdb_deconvoluter, s: *const mdb_fitting_settings) -> mdb_status;
    pub fn mdb_deconvoluter_add_ignore_region(d: *mut mdb_deconvoluter, lo: f64, hi: f64) -> mdb_status;
    pub fn mdb_deconvoluter_clear_ignore_regions(d: *mut mdb_deconvoluter);
    pub fn mdb_deconvoluter_set_superposition_mode(d: *mut mdb_deconvoluter, mode: c_int) -> mdb_status;
    pub fn mdb_deconvoluter_superposition_mode(d: *const mdb_deconvoluter) -> c_int;
    pub fn mdb_deconvoluter_set_fit_arithmetic(d: *mut mdb_deconvoluter, kind: c_int) -> mdb_status;
    pub fn mdb_deconvoluter_fit_arithmetic(d: *const mdb_deconvoluter) -> c_int;

    pub fn mdb_batch_len(b: *const mdb_batch) -> usize;
    pub fn mdb_batch_status(b: *const mdb_batch, i: usize) -> mdb_status;
    pub fn mdb_batch_n_lorentzians(b: *const mdb_batch, i: usize) -> usize;
    pub fn mdb_batch_lorentzians(b: *const mdb_batch, i: usize) -> *const mdb_lorentzian;
    pub fn mdb_batch_mse(b: *const mdb_batch, i: usize) -> f64;
    pub fn mdb_batch_n_peaks(b: *const mdb_batch, i: usize) -> usize;
    pub fn mdb_batch_peaks(b: *const mdb_batch, i: usize) -> *const i32;
    pub fn mdb_batch_free(b: *mut mdb_batch);
    pub fn mdb_batch_totals(b: *const mdb_batch, n_lorentzians: *mut usize, n_peaks: *mut usize);
    pub fn mdb_batch_export(
        b: *const mdb_batch,
        status: *mut i32,
        n_lorentzians: *mut u64,
        n_peaks: *mut u64,
        mse: *mut f64,
        lorentzians: *mut mdb_lorentzian,
        peaks: *mut i32,
    ) -> mdb_status;

    pub fn mdb_deconvolute_spectra(
        d: *const mdb_deconvoluter,
        spectra: *const mdb_spectrum_view,
        n_spectra: usize,
        memory: c_int,
        out: *mut *mut mdb_batch,
    ) -> mdb_status;
    pub fn mdb_deconvoluter_optimize_settings(
        d: *mut mdb_deconvoluter,
        reference: *const mdb_spectrum_view,
        memory: c_int,
        mse: *mut f64,
    ) -> mdb_status;
    pub fn mdb_superposition_vec(
        x: *const f64,
        n: usize,
        lorentzians: *const mdb_lorentzian,
        n_lorentzians: usize,
        out: *mut f64,
        memory: c_int,
    ) -> mdb_status;
    pub fn mdb_superposition_vec_mode(
        x: *const f64,
        n: usize,
        lorentzians: *const mdb_lorentzian,
        n_lorentzians: usize,
        out: *mut f64,
        memory: c_int,
        mode: c_int,
    ) -> mdb_status;

    pub fn mdb_stage_smooth(values: *const f64, n: usize, iterations: u64, window_size: u64, out: *mut f64) -> mdb_status;
    pub fn mdb_stage_smooth_batch(
        values_dev: *const f64,
        n: usize,
        count: usize,
        stride: usize,
        iterations: u64,
        window_size: u64,
        out_dev: *mut f64,
        ms: *mut f64,
    ) -> mdb_status;
    pub fn mdb_stage_detect(
        smoothed: *const f64,
        n: usize,
        peaks: *mut i32,
        scores: *mut f64,
        cap: usize,
        n_found: *mut usize,
    ) -> mdb_status;
    pub fn mdb_stage_select(
        d: *const mdb_deconvoluter,
        smoothed: *const f64,
        n: usize,
        sb0: usize,
        sb1: usize,
        has_ignore: c_int,
        ignore_idx: *const usize,
        n_ignore: usize,
        peaks: *mut i32,
        cap: usize,
        n_selected: *mut usize,
        mean_sd: *mut f64,
    ) -> mdb_status;
    pub fn mdb_stage_fit(
        x: *const f64,
        y: *const f64,
        n: usize,
        peaks: *const i32,
        n_peaks: usize,
        iterations: u64,
        out: *mut mdb_lorentzian,
        n_retained: *mut usize,
        trace: *mut mdb_lorentzian,
    ) -> mdb_status;
}
