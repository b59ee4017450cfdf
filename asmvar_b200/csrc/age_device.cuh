// age_device.cuh -- device-side task descriptors shared by the kernels and the host engine.
//
// Vocabulary (follows the reference, src/AsmvarAGE/src/AGEaligner.cpp):
//   aligner    one single-mode AGEaligner run (INDEL / TDUP / INVL / INVR): a forward fill, a reverse
//              fill, their running maxima, breakpoint enumeration and the traceback.
//   pair       two aligners with identical (len1, len2, scoring) that share one warp: aligner A lives in the
//              low int16 of every packed word, aligner B in the high int16 (s16x2 SIMD).
//   sweep      one direction (forward K1+K3 or reverse K2+K4) of a pair.
//   strip      256 consecutive contig rows = 32 lanes x 8 rows held in registers.
//   block      W = 4 consecutive reference-window columns: what one lane advances per wavefront step.
//   step       lane p works on block t-p in step t (anti-diagonal wavefront over lanes).
//   window     32 consecutive steps of one strip: checkpoint interval, tile of the split-maximum table and the
//              unit of trace recomputation.
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
constexpr int W      = 4;          // window columns per lane per step
constexpr int CKS    = 32;         // steps per window
constexpr int CKW    = 38;         // checkpoint words per lane: H,G,M of 8 rows + 3 x 4 outputs + 2 carried
constexpr int MAX_BP = 100;        // AGEaligner.hh:94

// trace codes, AGEaligner.hh:41-44
enum { T_STOP = 0, T_DIAG = 1, T_HORI = 2, T_VERT = 3 };

// mode flags, AGEaligner.hh:72-76
enum { F_INDEL = 0x01, F_TDUP = 0x02, F_INV = 0x04, F_INVR = 0x08, F_INVL = 0x10 };

struct PairTask {
    int32_t len1, len2, nstrips, P;      // P = nstrips * STRIP (padded rows)
    int32_t nb, nsteps, nwin;            // blocks = ceil(len1 / W), steps = nb + 31, windows = ceil(nsteps / CKS)
    int32_t flag[2];                     // single-mode flag of aligner A / B (0 = no aligner in that half)
    int32_t out[2];                      // index into the AlignerOut array (-1 = none)
    int32_t select;                      // 1: only the better-scoring half needs breakpoints (ties -> half 0):
                                         //    -both partners (age_align.cpp:273) or INVL/INVR of one -inv job (AGEaligner.cpp:151)
    uint32_t c0x, kabs, flip, gborder;   // packed per half; see the sweep
    uint32_t m2, mm2;                    // 2*match, 2*mismatch packed per half (special-column path)
    // inputs.  xblk[b] describes the 4 columns of logical block b (1..nb) in processing order:
    //   .x / .y = PRMT selectors of columns 0,1 / 2,3 (16 bit each); .z = per column byte: bit0 special, bits1-3 / 4-6 codes
    const uint4 *xblk_f, *xblk_r;        // [nb + 2]
    const uint2 *ytab;                   // [P] {Ta, Tb}: bytes = 2*S(x in ACGT, y_row) of aligner A / B
    const uint8_t *ycode;                // [P] low nibble = code of A's row, high nibble = B's
    // strip boundary rows, planes H, G, M: [(nstrips-1)][3][nb + 1] uint4 (4 columns of a block each)
    uint4 *bnd_f, *bnd_r;
    // Diagonal-major planes, [strip][step][...][lane]: what the 32 lanes touch in one step is contiguous, for the
    // forward and the reverse sweep alike (the anti-diagonal of one wavefront is the anti-diagonal of the other).
    uint32_t *D;                         // forward maxima shifted by one row and one column:
                                         //   [s][t][w][h][lane][4]: k = 4h+j -> 2*Fmax[256s + 8l + k][W(t-l-1) + w]
                                         //   = what the reverse cell (row+1, col+1) adds (K6, :545-550)
    uint32_t *Dlast;                     // [P + 1]  2*Fmax[r][len1]
    uint32_t *Dtail;                     // [W*nb + 1] 2*Fmax[P][c] (only read when len2 == P)
    uint32_t *Rfirst;                    // [P + 2]  2*Rmax[r][1]
    uint32_t *tilemax;                   // [nstrips][nwin][32] max over (lane's 8 rows x one window) of Fmax + Rmax
    uint32_t *ckpt_f, *ckpt_r;           // [nstrips][nwin][CKW][32] lane state at the start of every window
};

struct AlignerOut {
    int32_t status, score, n_bp, alt_index;   // n_bp = -1: score only (the half lost its selection)
    int32_t n_blk[4];                    // left, right, alt_left, alt_right
    int32_t blk_off[4];                  // offsets into the block pool (in blocks)
    int32_t bp[MAX_BP][4];
};

struct Block { int32_t start1, start2, length; };

// Per-warp scratch of pass 2: trace planes of the aligner being enumerated, filled window by window on demand
// by re-running the sweep from its checkpoints.
struct P2Scratch {
    uint2 *P[4];                         // 2-bit code planes FS, FM, RS, RM: [slots][32], 16 bits per block column
    uint32_t *hot;                       // [slots][32]: bit 8w+k = reverse cell holds Fmax + Rmax == max (with RM)
    uint32_t *Mp;                        // reverse maxima, same layout as D but unshifted (K7 only)
    uint32_t *valid;                     // [5][valid_words] window-valid bits: the four planes, Mp
    int32_t *rowrange;                   // [2][len2 + 2]
};

struct Pass2Params {
    const PairTask *pairs;
    int32_t n_pairs;
    AlignerOut *outs;
    Block *pool; int32_t pool_cap; int32_t *pool_used;
    int32_t *next;                       // work counter
    char *scratch; size_t scratch_stride; // per-warp scratch: warp w uses scratch + w * stride
    size_t max_slots, max_rows;          // layout of one warp's scratch
    int32_t need_mp;
};

// launchers (age_kernels.cu)
void launch_sweeps(const PairTask *d_pairs, int n_pairs, int32_t *d_counter, cudaStream_t st);
int  pass2_warps(int n_pairs);
size_t pass2_scratch_bytes(size_t max_slots, size_t max_rows, int need_mp);
void launch_pass2(const Pass2Params &p, int n_warps, cudaStream_t st);

} // namespace age
