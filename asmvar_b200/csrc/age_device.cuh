// age_device.cuh -- device-side task descriptors shared by the kernels and the host engine.
//
// Vocabulary (follows the reference, src/AsmvarAGE/src/AGEaligner.cpp):
//   aligner    one single-mode AGEaligner run (INDEL / TDUP / INVL / INVR): a forward fill, a reverse
//              fill, their running maxima, breakpoint enumeration and the traceback.
//   pair       two aligners with identical (len1, len2) that share one warp: aligner A lives in the
//              low int16 of every packed word, aligner B in the high int16 (s16x2 SIMD).
//   sweep      one direction (forward K1+K3 or reverse K2+K4) of a pair.
//   strip      256 consecutive contig rows = 32 lanes x 8 rows held in registers.
//
// Score representation inside the sweeps: every value is 2*score + tag in a signed 16-bit lane.
//   tag = 1 when the cell's own trace is HORIZONTAL/VERTICAL (AGEaligner.cpp:1006-1013: the gap
//   penalty for *leaving* a cell depends on that cell's trace).  max(2d, 2h+1, 2v+1, 0) then yields
//   the reference's tie order (gap beats diagonal on ties, negative clamps to 0 and clears the trace)
//   in one VIADDMNMX.S16x2.RELU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace age {

constexpr int RPT    = 8;          // contig rows per lane
constexpr int STRIP  = 32 * RPT;   // rows per strip
constexpr int MAX_BP = 100;        // AGEaligner.hh:94

// trace codes, AGEaligner.hh:41-44
enum { T_STOP = 0, T_DIAG = 1, T_HORI = 2, T_VERT = 3 };

// mode flags, AGEaligner.hh:72-76
enum { F_INDEL = 0x01, F_TDUP = 0x02, F_INV = 0x04, F_INVR = 0x08, F_INVL = 0x10 };

// xsel word: bits 0-15 PRMT selector, bit 16 = special column (N / U / other), bits 20-22 / 24-26 = codes
constexpr uint32_t XSEL_SPECIAL = 0x10000u;

struct PairTask {
    int32_t len1, len2, nstrips, P;      // P = nstrips * STRIP (padded rows)
    int32_t flag[2];                     // single-mode flag of aligner A / B (0 = no aligner in that half)
    int32_t out[2];                      // index into the AlignerOut array (-1 = none)
    uint32_t c0x, kabs, flip, gborder;   // packed per half; see sweep kernel
    uint32_t m2, mm2;                    // 2*match, 2*mismatch packed per half (special-column path)
    const uint32_t *xsel_f, *xsel_r;     // [len1 + 2], indexed by actual column c
    const uint2 *ytab;                   // [P] {Ta, Tb}: bytes = 2*S(x in ACGT, y_row) of aligner A / B
    const uint8_t *ycode;                // [P] low nibble = code of A's row, high nibble = B's
    uint4 *bnd_f, *bnd_r;                // strip boundary rows [(nstrips-1)][len1 + 2] {Hc, G2, M, xsel}
    int32_t select;                      // 1: only the better-scoring half needs breakpoints (ties -> half 0):
                                         //    -both partners (age_align.cpp:273) or INVL/INVR of one -inv job (AGEaligner.cpp:151)
    int32_t nwin;                        // 128-step windows per strip = ceil((len1 + 31) / 128)
    // The big planes are diagonal-major, [strip][step][lane][k]: what the 32 lanes of the wavefront touch in one
    // step is one contiguous 1 KB (maxima) or 256 B (trace) block, for the forward and for the reverse sweep alike
    // (the anti-diagonal of the forward wavefront is the anti-diagonal of the reverse one).
    uint32_t *D;                         // forward maxima shifted by one row: slot (s, t, l)[k] = 2*Fmax[256s+8l+k][t-l],
                                         //    which is what the reverse cell (256s+8l+k+1, t-l+1) adds (K6 :545-550)
    uint32_t *Dtail;                     // [len1 + 2] 2*Fmax[P][c] when len2 == P (row below the last lane)
    uint32_t *Mr;                        // reverse maxima: slot (s, t, l)[k] = 2*Rmax[256s+8l+k+1][len1+32-t-l]
    uint2 *Tf, *Tr;                      // trace nibbles (FS | FM << 2), one uint2 per slot: .x = aligner A, .y = B
    uint32_t *tilemax;                   // [nstrips][nwin][32] max over (lane's 8 rows x 128 steps) of Fmax + Rmax
    int32_t *rowrange;                   // [2][len2 + 2] pass-2 scratch: hot column range of every row
};

struct AlignerOut {
    int32_t status, score, n_bp, alt_index;   // n_bp = -1: score only (the half lost its selection)
    int32_t n_blk[4];                    // left, right, alt_left, alt_right
    int32_t blk_off[4];                  // offsets into the block pool (in blocks)
    int32_t bp[MAX_BP][4];
};

struct Block { int32_t start1, start2, length; };

struct Pass2Params {
    const PairTask *pairs;
    int32_t n_pairs;
    AlignerOut *outs;
    Block *pool; int32_t pool_cap; int32_t *pool_used;
    int32_t *next;                       // work counter
};

// launchers (age_kernels.cu)
void launch_sweeps(const PairTask *d_pairs, int n_pairs, int32_t *d_counter, cudaStream_t st);
void launch_pass2(const Pass2Params &p, cudaStream_t st);

} // namespace age
