/*
 * age_b200.h -- C ABI of the B200-native AGE (Alignment with Gap Excision) realigner.
 *
 * This is the drop-in boundary for the hot path of bioinformatics-centre/AsmVar's
 * src/AsmvarAGE.  The reference has no FFI of its own; the path sits behind the C++ class
 * AGEaligner and the age_align process.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference root, src/AsmvarAGE/src/...).
 *
 * Plain C, plain pointers and sizes, no C++ / torch types.  All CUDA work happens behind
 * this boundary in hand-written sm_100a kernels; there is no CPU fallback: if no CUDA
 * device is usable, age_create() fails and age_align_batch() returns an error.
 */
#ifndef AGE_B200_H
#define AGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Alignment modes -- the values of AGEaligner::*_FLAG (AGEaligner.hh:72-76). */
#define AGE_INDEL_FLAG        0x01
#define AGE_TDUPLICATION_FLAG 0x02
#define AGE_INVERSION_FLAG    0x04
#define AGE_INVR_FLAG         0x08
#define AGE_INVL_FLAG         0x10
#define AGE_MAX_BPOINTS       100   /* AGEaligner.hh:94 */

/* Status codes of age_result.status */
#define AGE_OK              0
#define AGE_NOT_ALIGNED     1   /* AGEaligner::align() would return false (AGEaligner.cpp:68-69) */
#define AGE_ERR_RANGE      -1   /* |match| or |mismatch| > 16383: 2*S does not fit the 16-bit score tables */
#define AGE_ERR_CUDA       -2
#define AGE_ERR_ARG        -3   /* also: several mode bits set without AGE_INVERSION_FLAG.  The reference accepts such flags
                                 * (AGEaligner.cpp:69 only tests `flag & ALL_FLAGS`) and then mixes the modes' code paths; neither of
                                 * its command lines can produce them.  With AGE_INVERSION_FLAG set the other bits are ignored, as in
                                 * AGEaligner.cpp:72-77.  Deliberate difference. */

typedef struct age_ctx age_ctx;

/* One AGEaligner(s1,s2) + align(scr,flag) call (AGEaligner.hh:101-103, AGEaligner.cpp:16-42,65-147;
 * scoring = Scorer(match,mismatch,gap_open,gap_extend), Scorer.hh:24-39).
 * seq1/seq2 are the nucleotide strings exactly as the reference's Sequence objects hold them
 * after substr()/revcom() (Sequence.hh:190-215); case is preserved, any byte < 0x80 is allowed. */
typedef struct {
    const char *seq1; int32_t len1;     /* first sequence  = target / reference window */
    const char *seq2; int32_t len2;     /* second sequence = query / assembly contig    */
    int32_t flag;                       /* one AGE_*_FLAG (see AGE_ERR_ARG)             */
    int16_t match, mismatch, gap_open, gap_extend;
    int32_t partner;                    /* -both (age_align.cpp:264-278): index in the same batch of the job that aligns the
                                         * reverse-complemented contig against the same seq1, or -1.  When two jobs name each other,
                                         * only the one age_align would print (the lower index if score >= the other's) gets
                                         * breakpoints and blocks; the other returns its score with detail = 0. */
} age_job;

/* A gap-free diagonal run; 1-based positions in seq1 / seq2 (class Block, AGEaligner.hh:128-177). */
typedef struct { int32_t start1, start2, length; } age_block;

/* Numbers of the printed alignment (bpoint 0), computed on the device from the block lists and the sequences, so that a
 * caller who wants figures and not text never has to rebuild the gapped strings:
 *   per fragment (0 = left, 1 = right) what AliFragment::countAligned counts (AGEaligner.hh:203-216): alignment columns,
 *   identical columns, columns with a gap;
 *   the runs printAlignment reports as "Identity at / outside / inside breakpoints" (AGEaligner.cpp:308-369, calcIdentity /
 *   calcOutsideIdentity / calcInsideIdentity :392-447) on the first and on the second sequence: run length, and for the "at"
 *   runs how far it reaches below (<= 0) and from the breakpoint on (>= 0).  All zero unless there are two fragments. */
typedef struct {
    int32_t n_aligned[2], n_identic[2], n_gap[2];
    int32_t homo_at1[3], homo_at2[3];
    int32_t homo_out1, homo_out2, homo_in1, homo_in2;
} age_counts;

/* Everything AGEaligner::score() / printAlignment() need (AGEaligner.cpp:149-154,179-380). */
typedef struct {
    int32_t status;
    int32_t detail;         /* 1: breakpoints and blocks are filled in; 0: score only (lost the -both selection) */
    int32_t score;          /* AGEaligner::score(): the aux (INVR) score only if strictly greater */
    int32_t flag;           /* single mode of the aligner whose record is printed (INVL/INVR for -inv) */
    int32_t used_aux;       /* 1 if the auxiliary INVR aligner is the printed one (AGEaligner.cpp:181-184) */
    int32_t main_score, aux_score;      /* _max of the main / auxiliary aligner (aux_score = -1 if none) */
    int32_t n_bp;                       /* _n_bpoints of the printed aligner */
    int32_t bp[AGE_MAX_BPOINTS][4];     /* _bpoints: lg1, lg2, rg1, rg2 (AGEaligner.cpp:658-661) */
    int32_t alt_index;                  /* bpoint shown under ALTERNATIVE REGION(S), -1 if none (AGEaligner.cpp:287-306) */
    int32_t n_left, n_right;            /* findLeftBlocks / findRightBlocks of bpoint 0 (AGEaligner.cpp:752-860) */
    int32_t n_alt_left, n_alt_right;    /* the same for bpoint alt_index */
    const age_block *left, *right, *alt_left, *alt_right;  /* owned by the ctx, valid until the next age_align_batch / age_stage call (age_submit: see there) */
    age_counts counts;                  /* of the alignment of bpoint 0 (left / right) */
} age_result;

/* A Sequence header: name, 1-based start coordinate and orientation (Sequence.hh:22-31). */
typedef struct {
    const char *seq; int32_t len;
    const char *name;
    int32_t start;
    int32_t reverse;
} age_seqview;

/* Device timing / accounting of the last age_align_batch call. */
typedef struct {
    double   kernel_ms;        /* CUDA-event time of all kernels of the batch (on the ctx stream) */
    double   fill_ms;          /* time of the DP sweep kernels alone */
    double   h2d_ms, d2h_ms;
    uint64_t cell_updates;     /* sum over executed aligners of 2*len1*len2 */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t launches;         /* kernels launched */
    uint64_t fill_launches;
    uint64_t n_aligners;       /* single-mode aligner runs executed (a -inv job is two) */
    uint64_t sweep_bytes;      /* bytes the sweep kernels of the batch move to and from HBM by construction: the D planes in the
                                * batch's storage format written and read back, checkpoints, strip boundary rows, tile maxima */
    uint64_t pool_reruns;      /* times (a device's part of) the batch was run again because its block pool or its trace pool
                                * was too small (the pools of later batches start larger) */
} age_stats;

/* Replaces: process start-up of age_align (age_align.cpp:194-229).  device_ids may be NULL
 * (= device 0); n_dev > 1 shards batches over several GPUs with no collective (host gather only). */
age_ctx *age_create(const int *device_ids, int n_dev);
void     age_destroy(age_ctx *ctx);
const char *age_last_error(const age_ctx *ctx);   /* ctx may be NULL: error of the failed age_create */
/* Allocate the scratch arena of every device now instead of with the first batches (a cudaMalloc of ~100 GB costs tens of
 * milliseconds, pinning a ~100 MB staging buffer even more): scratch_bytes per device; 0 = 80 % of what is free plus the
 * staging buffers of both batch slots at a size that fits batches of a few thousand candidates.  Optional; everything still
 * grows on demand. */
int age_reserve(age_ctx *ctx, size_t scratch_bytes);

/* Replaces: the serial candidate loop of age_align.cpp:231-293 / VarUnit.cpp:271-299 --
 * n independent AGEaligner::align() calls executed as length-binned GPU batches.
 * jobs and out are HOST arrays; returns 0 or a negative AGE_ERR_* (per-job status in out[i].status). */
int age_align_batch(age_ctx *ctx, const age_job *jobs, size_t n, age_result *out);

/* Streaming form of age_align_batch (the shape SURVEY 8(b) proposes for the C ABI): up to TWO batches may be in flight.
 * age_submit plans, packs and copies a batch and launches its kernels without waiting (jobs and the sequences they
 * point to may be released when it returns); age_collect waits for the OLDEST batch in flight and fills out[0..n-1]
 * (n = that batch's job count).  While the kernels of one batch run, the host prepares and uploads the next one and
 * the results of the previous one are read back.  The block pointers of a collected batch stay valid until the second
 * age_submit after its own.  age_align_batch(jobs, n, out) == age_submit + age_collect with nothing else in flight. */
int age_submit(age_ctx *ctx, const age_job *jobs, size_t n);
int age_collect(age_ctx *ctx, age_result *out, size_t n);
int age_in_flight(const age_ctx *ctx);

/* Split form of age_align_batch for measurement with inputs resident in HBM:
 * stage = host packing + host->device copy; run = kernels only (repeatable); fetch = device->host + result assembly. */
int age_stage(age_ctx *ctx, const age_job *jobs, size_t n);
int age_run_staged(age_ctx *ctx);
/* `reps` runs of the staged batch queued back to back without a host synchronisation in between (the chunks of consecutive
 * runs overlap exactly as the chunks of consecutive age_submit batches do); *total_ms = CUDA-event time of all of them. */
int age_run_staged_many(age_ctx *ctx, int reps, double *total_ms);
int age_fetch(age_ctx *ctx, age_result *out, size_t n);

int age_get_stats(const age_ctx *ctx, age_stats *st);

/* Host-only planning, no CUDA device needed (replaces nothing in the reference; its sharding hook is --selectId,
 * age_align.cpp:233): what age_create(ids, n_dev) + age_submit would do with these jobs.  device_of_job[i] = index of the
 * device the job's main aligner runs on (pairs are dealt longest-processing-time-first by DP area; mutual -both partners and
 * the INVL / INVR aligners of an -inv job share a warp, hence a device), -1 if the job is not aligned.
 * arith_of_job[i] (may be NULL) = 16: packed 16-bit halves, 32: wide 32-bit words, otherwise the AGE_* status (<= 1) the job would get. */
int age_plan(const age_job *jobs, size_t n, int n_dev, int32_t *device_of_job, int32_t *arith_of_job);

/* Replaces: AGEaligner::printAlignment() (AGEaligner.cpp:179-380) incl. AliFragment::printAlignment
 * (AGEaligner.hh:230-272).  Writes the record (starting with the blank line before "MATCH = ...") into
 * buf; returns the number of bytes needed (excluding the terminating NUL); if that is >= cap the output
 * was truncated.  s1/s2 are the headers of job->seq1/seq2. */
size_t age_format_alignment(const age_seqview *s1, const age_seqview *s2, const age_job *job,
                            const age_result *res, char *buf, size_t cap);

/* Replaces: AGEaligner::SetAlignResult() / align_result() of the AsmvarDetect flavour (src/AsmvarDetect/AGEaligner.cpp:179-332,
 * AGEaligner.h:91-150): the numbers printAlignment() prints, without going through text.  One age_fragment per aligned
 * fragment (MapData pair + AliFragment::countAligned); its gapped strings are written to the caller's `text` buffer as
 * three NUL-terminated strings of text_len characters each, starting at text_off: ali1, ali2, mapInfo ('|' '.' ' ').
 * Returns the number of text bytes needed (truncated if > cap).  When res->used_aux is set the reference's
 * SetAlignResult() prints the auxiliary record instead and leaves AlignResult at its defaults; that quirk is the
 * caller's (include/age_shim.hh reproduces it), this function describes whatever alignment `res` holds. */
typedef struct {
    int32_t start1, end1, start2, end2;
    int32_t n_aligned, n_identic, n_gap;
    int32_t text_off, text_len;
} age_fragment;
typedef struct {
    int32_t score;                      /* AlignResult::_score */
    int32_t strand;                     /* '+', '-' or '.' */
    int32_t is_alternative_align;       /* _n_bpoints > 1 (two fragments only) */
    int32_t n_identity;                 /* entries of _identity: total, then per fragment if there are two */
    int32_t identity[3][2];             /* (identical bases, percent) */
    int32_t ci_start1[2], ci_end1[2], ci_start2[2], ci_end2[2];
    int32_t homo_run_atbp1, homo_run_atbp2, homo_run_inbp1, homo_run_inbp2, homo_run_outbp1, homo_run_outbp2;
    int32_t n_frag;
    age_fragment frag[2];
} age_align_info;
size_t age_describe_alignment(const age_seqview *s1, const age_seqview *s2, const age_result *res,
                              age_align_info *info, char *text, size_t cap);

/* Replaces: AgeAlignment::VarReCall / CallVarInExcise of AsmvarDetect (src/AsmvarDetect/VarUnit.cpp:339-466) and the AGE=
 * sample field of the VCF (Variant.cpp:1259-1270) -- SURVEY 8(f)3.  From one age_result (its block lists and the counts the
 * device computed) and the headers of the two sequences, without going through text:
 *   identity      AlignResult::_identity: (identical bases, percent) of the whole alignment, then of either fragment when
 *                 there are two (AGEaligner.cpp:213-246 of the AsmvarDetect flavour)
 *   good_align    the good-alignment rule: average identity > 98 %, each flank > 30 identical bases and > 95 % (VarUnit.cpp:347-353)
 *   frag          the MapData of the fragments: start1, end1, start2, end2 in sequence coordinates
 *   ci_*          AlignResult::_ci_start1 ... _ci_end2: the span the breakpoints can slide over, from the three runs
 *   has_variant   two fragments, hence an excised region; then type / target_* / query_* / tlen / qlen are the variant
 *                 CallVarInExcise reports for it and cipos / ciend its confidence intervals relative to target_start / target_end
 * Returns 0, or AGE_ERR_ARG when res holds no alignment detail. */
#define AGE_VAR_NONE        0
#define AGE_VAR_DEL         1
#define AGE_VAR_INS         2
#define AGE_VAR_SNP         3
#define AGE_VAR_MNP         4
#define AGE_VAR_REPLACEMENT 5
typedef struct {
    int32_t n_identity, identity[3][2];
    int32_t strand;                     /* '+', '-' or '.' */
    int32_t good_align;
    int32_t n_frag, frag[2][4];
    int32_t ci_start1[2], ci_end1[2], ci_start2[2], ci_end2[2];
    int32_t homo_run;                   /* AlignResult::_homo_run_atbp1 */
    int32_t has_variant, type, tlen, qlen;
    int64_t target_start, target_end, query_start, query_end;
    int32_t cipos[2], ciend[2];
} age_variant;
int age_recall_variant(const age_seqview *s1, const age_seqview *s2, const age_result *res, age_variant *out);
const char *age_variant_type_name(int type);       /* "DEL", "INS", "SNP", "MNP", "REPLACEMENT" ("" for AGE_VAR_NONE) */
/* The value of the AGE= field: "T,+,<n>,<pct>,<n>,<pct>,<n>,<pct>" -- T/F = good_realign (the caller's isGoodReAlign: good_align
 * and a mis-mapping probability < 0.01, VarUnit.cpp:367), or "." when there is nothing to report (fewer than three
 * identity entries).  Returns the length needed. */
size_t age_format_vcf_age(const age_variant *v, int good_realign, char *buf, size_t cap);

/* Replaces: Sequence::revcom() string part (Sequence.hh:195-197) and Scorer lookup (Scorer.hh:24-39). */
void age_revcom(const char *in, int32_t n, char *out);
int  age_subst_score(char a, char b, int match, int mismatch);

/* Measurement helper (no counterpart in the reference): sustained issue rate of independent s16x2 SIMD chains
 * (VIMNMX / VIADDMNMX / VIMNMX3 / VIADD.16x2 / LOP3, the instruction mix of the DP recurrence) on every SM of
 * `device`, in 1e9 lane-instructions per second (one 32-bit SIMD instruction of one thread = 1).  This is the
 * denominator of bench.py's int_roofline.  Returns 0 or AGE_ERR_CUDA. */
int age_int16_peak(int device, double *giga_lane_instr_per_s);

const char *age_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGE_B200_H */
