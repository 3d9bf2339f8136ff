"""Error types mirroring lz4_flex's enums (src/block/mod.rs:82-106, src/frame/mod.rs:35-72)."""
from __future__ import annotations


class CompressError(Exception):
    """block::CompressError"""


class CompressOutputTooSmall(CompressError):
    """CompressError::OutputTooSmall"""

    def __str__(self):
        return ("output is too small for the compressed data, use get_maximum_output_size to "
                "reserve enough space")


class DecompressError(Exception):
    """block::DecompressError"""


class OutputTooSmall(DecompressError):
    """DecompressError::OutputTooSmall { expected, actual }"""

    def __init__(self, expected: int, actual: int):
        super().__init__(expected, actual)
        self.expected, self.actual = expected, actual

    def __str__(self):
        return (f"provided output is too small for the decompressed data, actual {self.actual}, "
                f"expected {self.expected}")


class LiteralOutOfBounds(DecompressError):
    def __str__(self):
        return "literal is out of bounds of the input"


class ExpectedAnotherByte(DecompressError):
    def __str__(self):
        return "expected another byte, found none"


class OffsetZero(DecompressError):
    def __str__(self):
        return "0 is not a valid match offset"


class OffsetOutOfBounds(DecompressError):
    def __str__(self):
        return "the offset to copy is not contained in the decompressed buffer"


class FrameError(Exception):
    """frame::Error"""


class CompressionError(FrameError):
    pass


class DecompressionError(FrameError):
    def __init__(self, inner: DecompressError):
        super().__init__(inner)
        self.inner = inner


class WrongMagicNumber(FrameError):
    pass


class ReservedBitsSet(FrameError):
    pass


class UnsupportedVersion(FrameError):
    pass


class UnsupportedBlocksize(FrameError):
    pass


class HeaderChecksumError(FrameError):
    pass


class BlockChecksumError(FrameError):
    pass


class ContentChecksumError(FrameError):
    pass


class ContentLengthError(FrameError):
    def __init__(self, expected=None, actual=None):
        super().__init__(expected, actual)
        self.expected, self.actual = expected, actual


class BlockTooBig(FrameError):
    pass


class SkippableFrame(FrameError):
    pass


class DictionaryNotSupported(FrameError):
    pass


class IoError(FrameError):
    pass


class LinkedBlocksUnsupported(FrameError):
    """BlockMode::Linked is not offered by the ENCODER (one serial chain per frame); the decoder handles it."""


class CudaError(RuntimeError):
    """The CUDA library or device is unavailable/failed.  There is no CPU fallback."""


_BLOCK = {3: LiteralOutOfBounds, 4: ExpectedAnotherByte, 5: OffsetZero, 6: OffsetOutOfBounds}
_FRAME = {
    102: WrongMagicNumber, 103: ReservedBitsSet, 104: UnsupportedVersion, 105: UnsupportedBlocksize,
    106: HeaderChecksumError, 107: BlockChecksumError, 108: ContentChecksumError, 109: ContentLengthError,
    110: BlockTooBig, 111: SkippableFrame, 112: DictionaryNotSupported, 113: IoError,
    114: LinkedBlocksUnsupported,
}


def block_error(status: int, expected: int = 0, actual: int = 0) -> Exception:
    if status == 1:
        return CompressOutputTooSmall()
    if status == 2:
        return OutputTooSmall(expected, actual)
    if status in _BLOCK:
        return _BLOCK[status]()
    return error_from_status(status)


def error_from_status(status: int, block_status: int = 0, detail: str = "", expected: int = 0,
                      actual: int = 0) -> Exception:
    if 1 <= status <= 6:
        return block_error(status, expected, actual)
    if status == 101:
        return DecompressionError(block_error(block_status, expected, actual))
    if status == 109:
        return ContentLengthError(expected, actual)
    if status in _FRAME:
        return _FRAME[status]()
    if status == 115:
        return IoError("output buffer exhausted")
    if status == 200:
        return ValueError("lz4b200: invalid argument")
    return CudaError(f"lz4b200: CUDA error {detail}".strip())
