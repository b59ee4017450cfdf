"""ctypes binding of the C ABI declared in include/age_b200.h.

The shared library (asmvar_b200/libage_b200.so) is built in-tree by ``__graft_entry__.build()``
(nvcc, sm_100a).  There is no Python or CPU fallback for the alignment itself: if the library
is missing, importing this module raises; if no CUDA device is usable, ``age_create`` fails.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("AGE_B200_LIB", PKG_DIR / "libage_b200.so"))

# mode flags (AGEaligner.hh:72-76)
INDEL_FLAG = 0x01
TDUPLICATION_FLAG = 0x02
INVERSION_FLAG = 0x04
INVR_FLAG = 0x08
INVL_FLAG = 0x10
MAX_BPOINTS = 100

AGE_OK = 0
AGE_NOT_ALIGNED = 1
AGE_ERR_RANGE = -1
AGE_ERR_CUDA = -2
AGE_ERR_ARG = -3


class AgeJob(C.Structure):
    _fields_ = [
        ("seq1", C.c_char_p), ("len1", C.c_int32),
        ("seq2", C.c_char_p), ("len2", C.c_int32),
        ("flag", C.c_int32),
        ("match", C.c_int16), ("mismatch", C.c_int16),
        ("gap_open", C.c_int16), ("gap_extend", C.c_int16),
        ("partner", C.c_int32),
    ]


class AgeBlock(C.Structure):
    _fields_ = [("start1", C.c_int32), ("start2", C.c_int32), ("length", C.c_int32)]


class AgeCounts(C.Structure):
    _fields_ = [("n_aligned", C.c_int32 * 2), ("n_identic", C.c_int32 * 2), ("n_gap", C.c_int32 * 2),
                ("homo_at1", C.c_int32 * 3), ("homo_at2", C.c_int32 * 3),
                ("homo_out1", C.c_int32), ("homo_out2", C.c_int32), ("homo_in1", C.c_int32), ("homo_in2", C.c_int32)]

    def as_dict(self) -> dict:
        return {"n_aligned": list(self.n_aligned), "n_identic": list(self.n_identic), "n_gap": list(self.n_gap),
                "homo_at1": list(self.homo_at1), "homo_at2": list(self.homo_at2),
                "homo_out": [self.homo_out1, self.homo_out2], "homo_in": [self.homo_in1, self.homo_in2]}


class AgeVariant(C.Structure):
    _fields_ = [("n_identity", C.c_int32), ("identity", (C.c_int32 * 2) * 3), ("strand", C.c_int32), ("good_align", C.c_int32),
                ("n_frag", C.c_int32), ("frag", (C.c_int32 * 4) * 2),
                ("ci_start1", C.c_int32 * 2), ("ci_end1", C.c_int32 * 2), ("ci_start2", C.c_int32 * 2), ("ci_end2", C.c_int32 * 2),
                ("homo_run", C.c_int32), ("has_variant", C.c_int32), ("type", C.c_int32), ("tlen", C.c_int32), ("qlen", C.c_int32),
                ("target_start", C.c_int64), ("target_end", C.c_int64), ("query_start", C.c_int64), ("query_end", C.c_int64),
                ("cipos", C.c_int32 * 2), ("ciend", C.c_int32 * 2)]

    def as_dict(self) -> dict:
        return {"identity": [list(self.identity[i]) for i in range(self.n_identity)], "strand": chr(self.strand), "good_align": bool(self.good_align),
                "frag": [list(self.frag[i]) for i in range(self.n_frag)],
                "ci_start1": list(self.ci_start1), "ci_end1": list(self.ci_end1), "ci_start2": list(self.ci_start2), "ci_end2": list(self.ci_end2),
                "homo_run": self.homo_run, "has_variant": bool(self.has_variant), "type": self.type, "tlen": self.tlen, "qlen": self.qlen,
                "target": [self.target_start, self.target_end], "query": [self.query_start, self.query_end],
                "cipos": list(self.cipos), "ciend": list(self.ciend)}


class AgeResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("detail", C.c_int32), ("score", C.c_int32), ("flag", C.c_int32), ("used_aux", C.c_int32),
        ("main_score", C.c_int32), ("aux_score", C.c_int32),
        ("n_bp", C.c_int32), ("bp", (C.c_int32 * 4) * MAX_BPOINTS),
        ("alt_index", C.c_int32),
        ("n_left", C.c_int32), ("n_right", C.c_int32),
        ("n_alt_left", C.c_int32), ("n_alt_right", C.c_int32),
        ("left", C.POINTER(AgeBlock)), ("right", C.POINTER(AgeBlock)),
        ("alt_left", C.POINTER(AgeBlock)), ("alt_right", C.POINTER(AgeBlock)),
        ("counts", AgeCounts),
    ]

    def bpoints(self):
        return [tuple(self.bp[i]) for i in range(self.n_bp)]

    def blocks(self, which: str):
        n = getattr(self, "n_" + which)
        p = getattr(self, which)
        return [(p[i].start1, p[i].start2, p[i].length) for i in range(n)]

    def summary(self) -> dict:
        """Plain-python view used by the parity tests."""
        return {
            "status": self.status, "detail": self.detail, "score": self.score, "flag": self.flag, "used_aux": self.used_aux,
            "main_score": self.main_score, "aux_score": self.aux_score,
            "bp": self.bpoints(), "alt_index": self.alt_index,
            "left": self.blocks("left"), "right": self.blocks("right"),
            "alt_left": self.blocks("alt_left") if self.alt_index >= 1 else [],
            "alt_right": self.blocks("alt_right") if self.alt_index >= 1 else [],
            "counts": self.counts.as_dict(),
        }


class AgeSeqView(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("len", C.c_int32), ("name", C.c_char_p),
                ("start", C.c_int32), ("reverse", C.c_int32)]


class AgeFragment(C.Structure):
    _fields_ = [("start1", C.c_int32), ("end1", C.c_int32), ("start2", C.c_int32), ("end2", C.c_int32),
                ("n_aligned", C.c_int32), ("n_identic", C.c_int32), ("n_gap", C.c_int32), ("text_off", C.c_int32), ("text_len", C.c_int32)]


class AgeAlignInfo(C.Structure):
    _fields_ = [("score", C.c_int32), ("strand", C.c_int32), ("is_alternative_align", C.c_int32), ("n_identity", C.c_int32),
                ("identity", (C.c_int32 * 2) * 3), ("ci_start1", C.c_int32 * 2), ("ci_end1", C.c_int32 * 2),
                ("ci_start2", C.c_int32 * 2), ("ci_end2", C.c_int32 * 2),
                ("homo_run_atbp1", C.c_int32), ("homo_run_atbp2", C.c_int32), ("homo_run_inbp1", C.c_int32),
                ("homo_run_inbp2", C.c_int32), ("homo_run_outbp1", C.c_int32), ("homo_run_outbp2", C.c_int32),
                ("n_frag", C.c_int32), ("frag", AgeFragment * 2)]


class AgeStats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("fill_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("cell_updates", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("launches", C.c_uint64), ("fill_launches", C.c_uint64), ("n_aligners", C.c_uint64), ("sweep_bytes", C.c_uint64), ("pool_reruns", C.c_uint64),
    ]


# every symbol include/age_b200.h declares; tests check the library exports all of them
EXPORTED_SYMBOLS = [
    "age_create", "age_destroy", "age_last_error", "age_align_batch", "age_stage", "age_run_staged", "age_run_staged_many",
    "age_fetch", "age_get_stats", "age_format_alignment", "age_revcom", "age_subst_score", "age_version", "age_int16_peak",
    "age_submit", "age_collect", "age_in_flight", "age_describe_alignment", "age_plan", "age_reserve", "age_recall_variant", "age_variant_type_name", "age_format_vcf_age",
]

_lib = None


def load() -> C.CDLL:
    """Load libage_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("AGE_LIB", LIB_PATH))      # AGE_LIB: another build of the same library (A/B measurements)
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    vp = C.c_void_p
    lib.age_create.restype = vp
    lib.age_create.argtypes = [C.POINTER(C.c_int), C.c_int]
    lib.age_destroy.restype = None
    lib.age_destroy.argtypes = [vp]
    lib.age_last_error.restype = C.c_char_p
    lib.age_last_error.argtypes = [vp]
    lib.age_align_batch.restype = C.c_int
    lib.age_align_batch.argtypes = [vp, C.POINTER(AgeJob), C.c_size_t, C.POINTER(AgeResult)]
    lib.age_submit.restype = C.c_int
    lib.age_submit.argtypes = [vp, C.POINTER(AgeJob), C.c_size_t]
    lib.age_collect.restype = C.c_int
    lib.age_collect.argtypes = [vp, C.POINTER(AgeResult), C.c_size_t]
    lib.age_in_flight.restype = C.c_int
    lib.age_in_flight.argtypes = [vp]
    lib.age_stage.restype = C.c_int
    lib.age_stage.argtypes = [vp, C.POINTER(AgeJob), C.c_size_t]
    lib.age_run_staged.restype = C.c_int
    lib.age_run_staged.argtypes = [vp]
    lib.age_run_staged_many.restype = C.c_int
    lib.age_run_staged_many.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    lib.age_fetch.restype = C.c_int
    lib.age_fetch.argtypes = [vp, C.POINTER(AgeResult), C.c_size_t]
    lib.age_get_stats.restype = C.c_int
    lib.age_get_stats.argtypes = [vp, C.POINTER(AgeStats)]
    lib.age_recall_variant.restype = C.c_int
    lib.age_recall_variant.argtypes = [C.POINTER(AgeSeqView), C.POINTER(AgeSeqView), C.POINTER(AgeResult), C.POINTER(AgeVariant)]
    lib.age_variant_type_name.restype = C.c_char_p
    lib.age_variant_type_name.argtypes = [C.c_int]
    lib.age_format_vcf_age.restype = C.c_size_t
    lib.age_format_vcf_age.argtypes = [C.POINTER(AgeVariant), C.c_int, C.c_char_p, C.c_size_t]
    lib.age_reserve.restype = C.c_int
    lib.age_reserve.argtypes = [vp, C.c_size_t]
    lib.age_plan.restype = C.c_int
    lib.age_plan.argtypes = [C.POINTER(AgeJob), C.c_size_t, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.age_format_alignment.restype = C.c_size_t
    lib.age_format_alignment.argtypes = [C.POINTER(AgeSeqView), C.POINTER(AgeSeqView), C.POINTER(AgeJob),
                                         C.POINTER(AgeResult), C.c_char_p, C.c_size_t]
    lib.age_describe_alignment.restype = C.c_size_t
    lib.age_describe_alignment.argtypes = [C.POINTER(AgeSeqView), C.POINTER(AgeSeqView), C.POINTER(AgeResult), C.POINTER(AgeAlignInfo),
                                           C.c_char_p, C.c_size_t]
    lib.age_revcom.restype = None
    lib.age_revcom.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
    lib.age_subst_score.restype = C.c_int
    lib.age_subst_score.argtypes = [C.c_char, C.c_char, C.c_int, C.c_int]
    lib.age_int16_peak.restype = C.c_int
    lib.age_int16_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.age_version.restype = C.c_char_p
    lib.age_version.argtypes = []
    _lib = lib
    return lib


def format_alignment(s1: AgeSeqView, s2: AgeSeqView, job: AgeJob, res: AgeResult) -> str:
    lib = load()
    cap = 4096 + 8 * (s1.len + s2.len)
    buf = C.create_string_buffer(cap)
    n = lib.age_format_alignment(C.byref(s1), C.byref(s2), C.byref(job), C.byref(res), buf, cap)
    if n >= cap:
        buf = C.create_string_buffer(n + 1)
        lib.age_format_alignment(C.byref(s1), C.byref(s2), C.byref(job), C.byref(res), buf, n + 1)
    return buf.raw[:n].decode("latin-1")


def recall_variant(s1: AgeSeqView, s2: AgeSeqView, res: AgeResult) -> dict:
    """age_recall_variant as a dict (None when res holds no alignment detail)"""
    lib = load()
    v = AgeVariant()
    if lib.age_recall_variant(C.byref(s1), C.byref(s2), C.byref(res), C.byref(v)) != 0:
        return None
    d = v.as_dict()
    d["type_name"] = lib.age_variant_type_name(v.type).decode()
    for good in (0, 1):
        buf = C.create_string_buffer(160)
        lib.age_format_vcf_age(C.byref(v), good, buf, 160)
        d["AGE=F" if not good else "AGE=T"] = buf.value.decode()
    return d


def describe_alignment(s1: AgeSeqView, s2: AgeSeqView, res: AgeResult) -> dict:
    """age_describe_alignment as a dict shaped like AsmvarDetect's AlignResult (src/AsmvarDetect/AGEaligner.h:91-150)."""
    lib = load()
    info = AgeAlignInfo()
    cap = 64 + 6 * (s1.len + s2.len)
    buf = C.create_string_buffer(cap)
    n = lib.age_describe_alignment(C.byref(s1), C.byref(s2), C.byref(res), C.byref(info), buf, cap)
    if n > cap:
        buf = C.create_string_buffer(n)
        lib.age_describe_alignment(C.byref(s1), C.byref(s2), C.byref(res), C.byref(info), buf, n)
    raw = buf.raw
    frags = []
    for k in range(info.n_frag):
        f = info.frag[k]
        t = [raw[f.text_off + i * (f.text_len + 1): f.text_off + i * (f.text_len + 1) + f.text_len].decode("latin-1") for i in range(3)]
        frags.append({"start1": f.start1, "end1": f.end1, "start2": f.start2, "end2": f.end2, "seq1": t[0], "seq2": t[1], "info": t[2],
                      "n_aligned": f.n_aligned, "n_identic": f.n_identic, "n_gap": f.n_gap})
    return {"score": info.score, "strand": chr(info.strand), "is_alternative_align": info.is_alternative_align,
            "identity": [[info.identity[i][0], info.identity[i][1]] for i in range(info.n_identity)],
            "ci_start1": list(info.ci_start1), "ci_end1": list(info.ci_end1), "ci_start2": list(info.ci_start2), "ci_end2": list(info.ci_end2),
            "homo_run": [info.homo_run_atbp1, info.homo_run_atbp2, info.homo_run_inbp1, info.homo_run_inbp2, info.homo_run_outbp1, info.homo_run_outbp2],
            "map": frags}
