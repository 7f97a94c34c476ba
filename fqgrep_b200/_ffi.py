"""ctypes binding of libfqgrep_b200.so (the C ABI in include/fqgrep_b200.h)."""
import ctypes as C
import os

from . import build as _build

_lib = None

SYMBOLS = [
    "fqg_last_error", "fqg_version", "fqg_matcher_new", "fqg_matcher_free", "fqg_validate_fixed_pattern",
    "fqg_matcher_dfa_info", "fqg_matcher_dfa_export", "fqg_ctx_create", "fqg_ctx_destroy", "fqg_device_count",
    "fqg_alloc_pinned", "fqg_free_pinned", "fqg_match_bases_device", "fqg_match_bases_host",
    "fqg_grep_fastq_host", "fqg_grep_fastq_device", "fqg_write_selected", "fqg_kernel_launches", "fqg_synth_reads_device", "fqg_fastq_split", "fqg_iupac_expand", "fqg_read_input", "fqg_free_input", "fqg_index_fastq_device", "fqg_compact_bases_device", "fqg_is_gzip_path", "fqg_is_fastq_path",
    "fqg_grep_fastq_host_write", "fqg_grep_fastq_host_multi", "fqg_stream_open", "fqg_stream_acquire", "fqg_stream_submit", "fqg_stream_feed",
    "fqg_stream_in_flight", "fqg_stream_next", "fqg_stream_close", "fqg_fastq_cut", "fqg_write_spans",
    "fqg_reader_open", "fqg_reader_read", "fqg_reader_close",
]


class DfaInfo(C.Structure):
    _fields_ = [("n_states", C.c_uint32), ("n_classes", C.c_uint32), ("start", C.c_uint32), ("accept_from", C.c_uint32),
                ("tier", C.c_uint32), ("table_bytes", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_units", C.c_uint64), ("n_selected", C.c_uint64),
                ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class Span(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("flags", C.c_uint32)]


class StreamOpts(C.Structure):
    _fields_ = [("n_slots", C.c_uint32), ("want_spans", C.c_uint32), ("slot_bytes", C.c_size_t)]


class Batch(C.Structure):
    _fields_ = [("seq", C.c_uint64), ("n_records", C.c_uint64), ("n_units", C.c_uint64), ("n_selected", C.c_uint64),
                ("unit_mask", C.POINTER(C.c_uint32)), ("spans", C.POINTER(Span)), ("n_spans", C.c_uint64),
                ("r1", C.c_void_p), ("r2", C.c_void_p), ("n1", C.c_size_t), ("n2", C.c_size_t),
                ("h2d_ms", C.c_double), ("kernel_ms", C.c_double)]


def lib():
    """Loads the in-tree shared library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FQG_LIB_VARIANT") or _build.SO      # measurement hook: a variant build of the same library (build.build_variant)
    if not os.path.exists(path):
        raise ImportError("libfqgrep_b200.so is not built: run `python -m fqgrep_b200.build` (nvcc, sm_100a). "
                          "fqgrep_b200 has no CPU or pure-Python path.")
    L = C.CDLL(path)
    vp, u64, u32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t
    L.fqg_last_error.restype = C.c_char_p
    L.fqg_version.restype = C.c_char_p
    L.fqg_kernel_launches.restype = u64
    L.fqg_matcher_new.argtypes = [C.c_int, C.c_char_p, vp, u32, u32, C.POINTER(vp)]
    L.fqg_matcher_free.argtypes = [vp]
    L.fqg_validate_fixed_pattern.argtypes = [C.c_char_p, C.c_int]
    L.fqg_matcher_dfa_info.argtypes = [vp, C.POINTER(DfaInfo)]
    L.fqg_matcher_dfa_export.argtypes = [vp, vp, vp]
    L.fqg_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.fqg_ctx_destroy.argtypes = [vp]
    L.fqg_alloc_pinned.argtypes = [sz, C.POINTER(vp)]
    L.fqg_free_pinned.argtypes = [vp]
    L.fqg_match_bases_device.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, u32, vp]
    L.fqg_match_bases_host.argtypes = [vp, vp, vp, vp, u64, vp, C.POINTER(u64)]
    L.fqg_grep_fastq_host.argtypes = [vp, vp, C.c_int, vp, sz, vp, sz, vp, sz, C.POINTER(Result)]
    L.fqg_grep_fastq_device.argtypes = [vp, vp, C.c_int, vp, sz, vp, sz, vp, sz, C.POINTER(Result), vp]
    L.fqg_write_selected.argtypes = [C.c_int, vp, sz, vp, sz, vp, vp, C.POINTER(sz), vp, C.POINTER(sz)]
    L.fqg_synth_reads_device.argtypes = [vp, vp, u64, u64, u32, C.c_int, u64, C.c_char_p, vp, u32, C.c_double, u64, vp]
    L.fqg_fastq_split.argtypes = [vp, sz, u32, vp]
    L.fqg_index_fastq_device.argtypes = [vp, vp, sz, vp, vp, u64, C.POINTER(u64), vp]
    L.fqg_compact_bases_device.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp]
    L.fqg_read_input.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.fqg_free_input.argtypes = [vp, C.c_int]
    L.fqg_is_gzip_path.argtypes = [C.c_char_p]
    L.fqg_is_fastq_path.argtypes = [C.c_char_p]
    L.fqg_iupac_expand.argtypes = [C.c_char_p, C.c_int, vp, sz, C.POINTER(sz)]
    L.fqg_grep_fastq_host_write.argtypes = [vp, vp, C.c_int, vp, sz, vp, sz, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz), C.POINTER(Result)]
    L.fqg_grep_fastq_host_multi.argtypes = [C.POINTER(vp), u32, vp, C.c_int, vp, sz, vp, sz, vp, sz, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz),
                                            C.POINTER(Result)]
    L.fqg_stream_open.argtypes = [vp, vp, C.c_int, C.POINTER(StreamOpts), C.POINTER(vp)]
    L.fqg_stream_acquire.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
    L.fqg_stream_submit.argtypes = [vp, sz, sz]
    L.fqg_stream_feed.argtypes = [vp, vp, sz, vp, sz]
    L.fqg_stream_in_flight.argtypes = [vp]
    L.fqg_stream_next.argtypes = [vp, C.POINTER(Batch)]
    L.fqg_stream_close.argtypes = [vp]
    L.fqg_stream_close.restype = None
    L.fqg_reader_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.fqg_reader_read.argtypes = [vp, vp, sz, C.POINTER(sz), C.POINTER(C.c_int)]
    L.fqg_reader_close.argtypes = [vp]
    L.fqg_reader_close.restype = None
    L.fqg_fastq_cut.argtypes = [C.c_int, vp, sz, vp, sz, sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    L.fqg_write_spans.argtypes = [C.POINTER(Span), u64, vp, sz, vp, sz, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]
    _lib = L
    return L


def last_error():
    return lib().fqg_last_error().decode("utf-8", "replace")
