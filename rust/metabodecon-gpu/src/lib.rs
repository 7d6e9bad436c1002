//! Safe wrapper over `metabodecon-sys`.  Method names and semantics follow
//! `metabodecon::deconvolution::Deconvoluter` (metabodecon/src/deconvolution/deconvoluter.rs) and
//! `Lorentzian::superposition_vec` (lorentzian.rs:631-635); inputs are plain slices so that the
//! `metabodecon` crate can adapt its own `Spectrum` / `Deconvolution` types behind a `gpu` feature
//!.  There is no CPU fallback: every call needs a CUDA device.
use metabodecon_sys as sys;
use std::ffi::CStr;
use std::ptr;

/// Mirrors `metabodecon::deconvolution::error::Kind` plus the transport-level failures.
#[derive(Debug, Clone, PartialEq)]
pub enum Error {
    NoPeaksDetected,
    EmptySignalRegion,
    EmptySignalFreeRegion,
    InvalidSmoothingSettings,
    InvalidSelectionSettings,
    InvalidFittingSettings,
    InvalidIgnoreRegion,
    /// input on which the reference implementation panics
    ReferencePanic(String),
    /// CUDA failure or no device (message from `mdb_last_error_message`)
    Cuda(String),
    Other(i32, String),
}

pub type Result<T> = std::result::Result<T, Error>;

fn last_message() -> String {
    unsafe {
        let p = sys::mdb_last_error_message();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

fn check(status: sys::mdb_status) -> Result<()> {
    match status {
        sys::MDB_OK => Ok(()),
        sys::MDB_ERR_NO_PEAKS_DETECTED => Err(Error::NoPeaksDetected),
        sys::MDB_ERR_EMPTY_SIGNAL_REGION => Err(Error::EmptySignalRegion),
        sys::MDB_ERR_EMPTY_SIGNAL_FREE_REGION => Err(Error::EmptySignalFreeRegion),
        sys::MDB_ERR_INVALID_SMOOTHING_SETTINGS => Err(Error::InvalidSmoothingSettings),
        sys::MDB_ERR_INVALID_SELECTION_SETTINGS => Err(Error::InvalidSelectionSettings),
        sys::MDB_ERR_INVALID_FITTING_SETTINGS => Err(Error::InvalidFittingSettings),
        sys::MDB_ERR_INVALID_IGNORE_REGION => Err(Error::InvalidIgnoreRegion),
        sys::MDB_ERR_REFERENCE_PANIC => Err(Error::ReferencePanic(last_message())),
        sys::MDB_ERR_CUDA => Err(Error::Cuda(last_message())),
        other => Err(Error::Other(other, last_message())),
    }
}

/// `Lorentzian {sfhw, hw2, maxp}`; layout-compatible with `mdb_lorentzian`.
pub type Lorentzian = sys::mdb_lorentzian;

/// What the path reads of a `Spectrum`: both axes and the (axis-ordered) signal boundaries.
#[derive(Clone, Copy)]
pub struct SpectrumRef<'a> {
    pub chemical_shifts: &'a [f64],
    pub intensities: &'a [f64],
    pub signal_boundaries: (f64, f64),
}

/// One deconvoluted spectrum (`Deconvolution`, deconvolution.rs:45-56, without the settings copy).
#[derive(Debug, Clone)]
pub struct Deconvolution {
    pub lorentzians: Vec<Lorentzian>,
    pub mse: f64,
}

pub struct Deconvoluter {
    handle: *mut sys::mdb_deconvoluter,
}

// The C handle is immutable during deconvolution calls and the library is re-entrant
// (deconvoluter.rs:913-917 asserts the same for the reference type).
unsafe impl Send for Deconvoluter {}
unsafe impl Sync for Deconvoluter {}

impl Deconvoluter {
    /// `Deconvoluter::default()`: MovingAverage(3, 3), NoiseScoreFilter(MinimumSum, 5.0), Analytical(10).
    pub fn new_default() -> Result<Self> {
        let mut handle = ptr::null_mut();
        check(unsafe { sys::mdb_deconvoluter_default(&mut handle) })?;
        Ok(Self { handle })
    }

    pub fn set_moving_average_smoother(&mut self, iterations: usize, window_size: usize) -> Result<()> {
        let s = sys::mdb_smoothing_settings {
            kind: sys::MDB_SMOOTHING_MOVING_AVERAGE,
            iterations: iterations as u64,
            window_size: window_size as u64,
        };
        check(unsafe { sys::mdb_deconvoluter_set_smoothing_settings(self.handle, &s) })
    }

    pub fn set_noise_score_selector(&mut self, threshold: f64) -> Result<()> {
        let s = sys::mdb_selection_settings {
            kind: sys::MDB_SELECTION_NOISE_SCORE_FILTER,
            scoring_method: sys::MDB_SCORING_MINIMUM_SUM,
            threshold,
        };
        check(unsafe { sys::mdb_deconvoluter_set_selection_settings(self.handle, &s) })
    }

    pub fn set_analytical_fitter(&mut self, iterations: usize) -> Result<()> {
        let s = sys::mdb_fitting_settings { kind: sys::MDB_FITTING_ANALYTICAL, iterations: iterations as u64 };
        check(unsafe { sys::mdb_deconvoluter_set_fitting_settings(self.handle, &s) })
    }

    pub fn add_ignore_region(&mut self, region: (f64, f64)) -> Result<()> {
        check(unsafe { sys::mdb_deconvoluter_add_ignore_region(self.handle, region.0, region.1) })
    }

    pub fn clear_ignore_regions(&mut self) {
        unsafe { sys::mdb_deconvoluter_clear_ignore_regions(self.handle) }
    }

    fn view(s: &SpectrumRef<'_>) -> sys::mdb_spectrum_view {
        assert_eq!(s.chemical_shifts.len(), s.intensities.len());
        sys::mdb_spectrum_view {
            chemical_shifts: s.chemical_shifts.as_ptr(),
            intensities: s.intensities.as_ptr(),
            len: s.intensities.len(),
            signal_boundaries: [s.signal_boundaries.0, s.signal_boundaries.1],
        }
    }

    /// `deconvolute_spectra` / `par_deconvolute_spectra` (deconvoluter.rs:651-661, 699-710): the
    /// first failing spectrum in index order decides the error.
    pub fn deconvolute_spectra(&self, spectra: &[SpectrumRef<'_>]) -> Result<Vec<Deconvolution>> {
        let views: Vec<_> = spectra.iter().map(Self::view).collect();
        let mut batch = ptr::null_mut();
        let status = unsafe {
            sys::mdb_deconvolute_spectra(self.handle, views.as_ptr(), views.len(), sys::MDB_MEM_HOST, &mut batch)
        };
        struct Guard(*mut sys::mdb_batch);
        impl Drop for Guard {
            fn drop(&mut self) {
                if !self.0.is_null() {
                    unsafe { sys::mdb_batch_free(self.0) }
                }
            }
        }
        let _guard = Guard(batch);
        check(status)?;
        Ok((0..views.len())
            .map(|i| unsafe {
                let n = sys::mdb_batch_n_lorentzians(batch, i);
                let p = sys::mdb_batch_lorentzians(batch, i);
                let lorentzians = if n == 0 { Vec::new() } else { std::slice::from_raw_parts(p, n).to_vec() };
                Deconvolution { lorentzians, mse: sys::mdb_batch_mse(batch, i) }
            })
            .collect())
    }

    pub fn par_deconvolute_spectra(&self, spectra: &[SpectrumRef<'_>]) -> Result<Vec<Deconvolution>> {
        self.deconvolute_spectra(spectra)
    }

    /// `deconvolute_spectrum` / `par_deconvolute_spectrum` (deconvoluter.rs:530-552, 590-613).
    pub fn deconvolute_spectrum(&self, spectrum: &SpectrumRef<'_>) -> Result<Deconvolution> {
        Ok(self.deconvolute_spectra(std::slice::from_ref(spectrum))?.remove(0))
    }

    pub fn par_deconvolute_spectrum(&self, spectrum: &SpectrumRef<'_>) -> Result<Deconvolution> {
        self.deconvolute_spectrum(spectrum)
    }

    /// `optimize_settings` (deconvoluter.rs:761-825): returns the lowest MSE and keeps its settings.
    pub fn optimize_settings(&mut self, reference: &SpectrumRef<'_>) -> Result<f64> {
        let view = Self::view(reference);
        let mut mse = 0.0;
        check(unsafe { sys::mdb_deconvoluter_optimize_settings(self.handle, &view, sys::MDB_MEM_HOST, &mut mse) })?;
        Ok(mse)
    }
}

impl Clone for Deconvoluter {
    fn clone(&self) -> Self {
        let mut handle = ptr::null_mut();
        let status = unsafe { sys::mdb_deconvoluter_clone(self.handle, &mut handle) };
        assert_eq!(status, sys::MDB_OK, "mdb_deconvoluter_clone failed");
        Self { handle }
    }
}

impl Deconvoluter {
    /// Pins the arithmetic of `Deconvolution::mse` for THIS deconvoluter (`mdb_deconvoluter_set_superposition_mode`):
    /// two deconvoluters of one process may differ and run concurrently.  Unpinned ones follow
    /// `set_exact_superposition`'s process default.
    pub fn set_exact_mse(&mut self, exact: bool) -> Result<()> {
        check(unsafe {
            sys::mdb_deconvoluter_set_superposition_mode(
                self.handle,
                if exact { sys::MDB_SUPERPOSITION_EXACT } else { sys::MDB_SUPERPOSITION_FAST },
            )
        })
    }
}

impl Drop for Deconvoluter {
    fn drop(&mut self) {
        unsafe { sys::mdb_deconvoluter_free(self.handle) }
    }
}

/// `Lorentzian::superposition_vec` / `par_superposition_vec` (lorentzian.rs:631-635, 656-663).
pub fn superposition_vec(x: &[f64], lorentzians: &[Lorentzian]) -> Result<Vec<f64>> {
    let mut out = vec![0.0; x.len()];
    check(unsafe {
        sys::mdb_superposition_vec(x.as_ptr(), x.len(), lorentzians.as_ptr(), lorentzians.len(), out.as_mut_ptr(), sys::MDB_MEM_HOST)
    })?;
    Ok(out)
}

/// `superposition_vec` with the arithmetic stated by the caller instead of the process default.
pub fn superposition_vec_with(x: &[f64], lorentzians: &[Lorentzian], exact: bool) -> Result<Vec<f64>> {
    let mut out = vec![0.0; x.len()];
    let mode = if exact { sys::MDB_SUPERPOSITION_EXACT } else { sys::MDB_SUPERPOSITION_FAST };
    check(unsafe {
        sys::mdb_superposition_vec_mode(x.as_ptr(), x.len(), lorentzians.as_ptr(), lorentzians.len(), out.as_mut_ptr(), sys::MDB_MEM_HOST, mode)
    })?;
    Ok(out)
}

/// Process DEFAULT of the arithmetic of `Deconvolution::mse` and `superposition_vec` (include/mdb200.h,
/// `mdb_set_superposition_mode`): `true` replays the reference's operators bit for bit, `false`
/// (the library default) uses half the FP64 instructions and agrees to about 1e-15 relative (1e-13 for the MSE).
/// Peak sets and Lorentzian parameters are bit-identical in both.
pub fn set_exact_superposition(exact: bool) -> Result<()> {
    check(unsafe {
        sys::mdb_set_superposition_mode(if exact { sys::MDB_SUPERPOSITION_EXACT } else { sys::MDB_SUPERPOSITION_FAST })
    })
}
