"""Exception tree of the reference's Python package (metabodecon-python/src/error.rs:5-28)."""


class Error(Exception):
    """Base class for all metabodecon errors."""


class UnexpectedError(Error):
    """An unexpected error occurred (includes CUDA failures and inputs the reference panics on)."""


class ThreadPoolError(Error):
    """Kept for API compatibility; the GPU path has no thread pool."""


class SerializationError(Error):
    """Serialization or deserialization failed."""


class SpectrumError(Error):
    """Errors of the Spectrum class."""


class EmptyData(SpectrumError):
    """Input data is empty."""


class DataLengthMismatch(SpectrumError):
    """Input data lengths do not match."""


class NonUniformSpacing(SpectrumError):
    """Chemical shifts are not uniformly spaced."""


class InvalidIntensities(SpectrumError):
    """Intensities contain invalid values."""


class InvalidSignalBoundaries(SpectrumError):
    """Signal boundaries are invalid."""


class MissingMetadata(SpectrumError):
    """Metadata is missing from a spectrum file."""


class MalformedMetadata(SpectrumError):
    """Metadata in a spectrum file is malformed."""


class MissingData(SpectrumError):
    """Data is missing from a spectrum file."""


class MalformedData(SpectrumError):
    """Data in a spectrum file is malformed."""


class DeconvolutionError(Error):
    """Errors of the deconvolution process."""


class InvalidSmoothingSettings(DeconvolutionError):
    """Smoothing settings are invalid."""


class InvalidSelectionSettings(DeconvolutionError):
    """Selection settings are invalid."""


class InvalidFittingSettings(DeconvolutionError):
    """Fitting settings are invalid."""


class InvalidIgnoreRegion(DeconvolutionError):
    """Ignore region is invalid."""


class NoPeaksDetected(DeconvolutionError):
    """No peaks were detected in the spectrum."""


class EmptySignalRegion(DeconvolutionError):
    """No peaks were found in the signal region."""


class EmptySignalFreeRegion(DeconvolutionError):
    """No peaks were found in the signal free region."""


class CudaError(UnexpectedError):
    """The CUDA device or runtime failed, or no device is present (there is no CPU fallback)."""


# mdb_status -> exception class (metabodecon-python/src/error.rs:46-96)
_BY_STATUS = {
    1: NoPeaksDetected,
    2: EmptySignalRegion,
    3: EmptySignalFreeRegion,
    4: InvalidSmoothingSettings,
    5: InvalidSelectionSettings,
    6: InvalidFittingSettings,
    7: InvalidIgnoreRegion,
    10: EmptyData,
    11: DataLengthMismatch,
    12: NonUniformSpacing,
    13: InvalidIntensities,
    14: InvalidSignalBoundaries,
    100: UnexpectedError,
    200: CudaError,
    201: UnexpectedError,
    202: UnexpectedError,
}


def raise_for_status(status: int, message: str = "") -> None:
    if status == 0:
        return
    raise _BY_STATUS.get(status, UnexpectedError)(message or f"mdb_status {status}")
