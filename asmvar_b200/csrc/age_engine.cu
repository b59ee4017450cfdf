// age_engine.cu -- host side of the C ABI declared in include/age_b200.h.
//
// Replaces the serial candidate loop of the reference (age_align.cpp:231-293, VarUnit.cpp:271-299) with
// batched GPU execution: jobs are validated, expanded into single-mode aligners (-inv = INVL + auxiliary
// INVR, AGEaligner.cpp:72-77,136-144), paired by shape, packed, swept and enumerated on the device(s), and
// the compact results (score, breakpoints, diagonal blocks) are gathered on the host in input order.
// Two batches can be in flight (age_submit / age_collect): while the kernels of one run, the next one is planned,
// packed and copied to the device on a second stream, and the results of the previous one are read back.
// There is no CPU implementation of the alignment here: without a usable CUDA device age_create() fails.
#include "age_b200.h"
#include "age_device.cuh"
#include "age_host.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstddef>
#include <string>
#include <thread>
#include <vector>

using namespace age;

namespace {

thread_local std::string g_create_error;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const bool g_hprof = getenv("AGE_HOST_PROF") != nullptr;

// Device / pinned arenas grow with 1/16 of headroom so that a stream of similar batches does not reallocate
// (cudaMalloc of a ~100 GB arena costs a few hundred milliseconds).
struct Arena {
    char *ptr = nullptr; size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        const double t0 = wall_ms();
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        size_t want = n + n / 16;
        if (cudaMalloc(&ptr, want) != cudaSuccess) {
            cudaGetLastError(); want = n;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        cap = want;
        if (g_hprof) fprintf(stderr, "[host prof] cudaMalloc %.1f MB: %.1f ms\n", want / 1e6, wall_ms() - t0);
        return true;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};
struct PinnedBuf {
    char *ptr = nullptr; size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        const double t0 = wall_ms();
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr; cap = 0;
        size_t want = n + n / 16;
        if (cudaMallocHost(&ptr, want) != cudaSuccess) { cudaGetLastError(); return false; }
        cap = want;
        if (g_hprof) fprintf(stderr, "[host prof] cudaMallocHost %.1f MB: %.1f ms\n", want / 1e6, wall_ms() - t0);
        return true;
    }
    void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
};

struct Aligner {                 // one single-mode AGEaligner run
    int job, role;               // role 0 = main, 1 = auxiliary INVR aligner of an -inv job
    int flag;
    int out;                     // index into the AlignerOut array of its sub-batch
};

struct HostPair {
    int a[2];                    // aligner indices (second may be -1)
    int select;                  // 1: only the better half needs breakpoints (see PairTask::select)
    int len1, len2, nstrips, P, nb, nsteps, nwin;
    int k7;                      // holds a TDUP / INVL / INVR aligner (pass 2 then needs the reverse maxima plane)
    int dfmt;                    // storage format of the D plane (DFMT_RAW / DFMT_OFF8, age_device.cuh)
    int wide;                    // one aligner in 32-bit words (PairTask::wide)
    int nsets; size_t set_bytes;  // pass-2 page tables and lists (one set per aligner that gets breakpoints)
    size_t trace_share;          // this pair's contribution to the batch's trace pool, in granules (age_device.cuh: P2Scratch)
    size_t trace_full;           // granules all trace pages of the pair would take
    size_t in_bytes, scratch_bytes;
    double cells;
};

// What one batch in flight owns on one device.  The big scratch arena (D planes, checkpoints, trace planes) is
// shared by both slots of a device: kernels of consecutive batches run back to back on one stream.
struct Slot {
    Arena in, outbuf, aux;
    PinnedBuf hin, hout;
    cudaEvent_t ev[7] = {};      // 0/1 H2D begin/end (cp), 2 kernels begin, 3 sweeps end, 4 kernels end (st), 5/6 D2H begin/end (rb)
    std::vector<int> pair_ids;   // the sub-batch resident in this slot: wide pairs first, then longest first
    std::vector<PairTask> tasks;
    std::vector<size_t> sc_off;
    int n_outs = 0, pool_cap = 0, queue_cap = 0, n_wide = 0, nw = 1, nw_wide = 1, cs = 1, cs_wide = 1, launches = 0;
    size_t in_bytes = 0, out_bytes = 0, sc_bytes = 0;
    size_t tpool_off = 0, tpool_full = 0; uint32_t tpool_granules = 0;   // trace pool of the sub-batch: behind the pairs' scratch
};

struct Device {
    int id = 0;
    cudaStream_t st = nullptr;   // kernels
    cudaStream_t cp = nullptr;   // host -> device copies of the next batch
    cudaStream_t rb = nullptr;   // device -> host read-back of the previous one
    Arena scratch;
    Slot slot[2];
    double trace_scale = 1.0;    // grows when a batch overflowed its trace pool: later batches of the same kind get more from the start
    std::string err;
};

// The plan and the gathered results of one batch
struct Batch {
    size_t n_jobs = 0;
    std::vector<int> job_status;
    std::vector<Aligner> aligners;
    std::vector<HostPair> pairs;
    std::vector<AlignerOut> outs;                 // per aligner (global index)
    std::vector<std::vector<age_block>> blocks;   // per aligner: left | right | alt_left | alt_right
    std::vector<int> partner_of;                  // mutual -both partner of every job (or -1)
    std::vector<char> on_dev;                     // devices that hold an asynchronous sub-batch of this batch
    age_stats stats{};
    int state = 0;                                // 0 free, 1 in flight on the devices, 2 results gathered (oversize batch, run synchronously)
};

} // namespace

struct age_ctx {
    std::vector<Device> devs;
    std::string err;
    age_stats stats{};                            // of the batch collected last
    Batch batch[2];
    int head = 0;                                 // slot of the next age_submit
    int in_flight = 0;
    bool staged = false;
    int host_threads = 1;
};

namespace {

template <class F> void parallel_for(int nthreads, size_t n, F f)
{
    if (nthreads <= 1 || n < 64) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&] { for (;;) { size_t i = next.fetch_add(16); if (i >= n) break; for (size_t k = i; k < std::min(n, i + 16); ++k) f(k); } });
    for (auto &x : th) x.join();
}

bool single_mode(int flag) { return flag == F_INDEL || flag == F_TDUP || flag == F_INVL || flag == F_INVR; }

// Which arithmetic a job runs in (DESIGN.md "Arithmetic and its limits").  PACKED: 2*score + tag fits a signed 16-bit half
// for every cell, two aligners share a word.  WIDE: 32-bit words with the reference's unsigned-short storage made explicit
// (AGEaligner.hh:86): any gap costs, any length, scores beyond 32767 and the wrap beyond 65535 included.
enum { RANGE_PACKED = 0, RANGE_WIDE = 1, RANGE_ERR = 2 };
int classify_range(const age_job &j, bool force_wide)
{
    const int m = j.match, mm = j.mismatch, go = j.gap_open, ge = j.gap_extend;
    if (m > 16383 || m < -16383 || mm > 16383 || mm < -16383) return RANGE_ERR;  // 2*S must fit the 16-bit table entries of the wide path
    if (force_wide) return RANGE_WIDE;
    if (m > 63 || m < -63 || mm > 63 || mm < -63) return RANGE_WIDE;
    if (go >= 0 || ge >= 0 || go < -4000 || ge < -4000) return RANGE_WIDE;
    const long k = 2L * (ge - go) - 1;
    const long kabs = k < 0 ? -k : k;
    const long top = 2L * std::max(std::max(m, mm), 0) * std::min(j.len1, j.len2);   // a DP step adds at most max(match, mismatch, 0)
    if (top + kabs + 2L * (-go) + 2L * (-ge) + 8 > 32767) return RANGE_WIDE;
    return RANGE_PACKED;
}

void build_plan(Batch *B, const age_job *jobs, size_t n)
{
    const char *dfmt_env = getenv("AGE_DFMT");                      // test knob: "raw" keeps the 16-bit D planes
    const bool force_raw = dfmt_env && strcmp(dfmt_env, "raw") == 0;
    B->n_jobs = n;
    B->job_status.assign(n, AGE_OK);
    B->aligners.clear(); B->pairs.clear();
    std::vector<char> wide_job(n, 0);
    const bool force_wide = getenv("AGE_FORCE_WIDE") != nullptr;    // test knob: everything through the 32-bit kernels
    for (size_t i = 0; i < n; ++i) {
        const age_job &j = jobs[i];
        int st = AGE_OK;
        // AGEaligner.cpp:72-77: the INVERSION bit wins over whatever else is set (main aligner INVL, auxiliary INVR)
        const bool inv = (j.flag & F_INV) != 0;
        if (!j.seq1 || !j.seq2 || j.len1 <= 0 || j.len2 <= 0) st = AGE_NOT_ALIGNED;        // AGEaligner.cpp:68
        else if (!(j.flag & (F_INDEL | F_TDUP | F_INV | F_INVR | F_INVL))) st = AGE_NOT_ALIGNED;   // AGEaligner.cpp:69
        else if (!inv && !single_mode(j.flag)) st = AGE_ERR_ARG;        // several mode bits without INVERSION: see age_b200.h
        else {
            const int rg = classify_range(j, force_wide);
            if (rg == RANGE_ERR) st = AGE_ERR_RANGE;
            wide_job[i] = rg == RANGE_WIDE;
        }
        B->job_status[i] = st;
        if (st != AGE_OK) continue;
        if (inv) {
            B->aligners.push_back({(int)i, 0, F_INVL, -1});
            B->aligners.push_back({(int)i, 1, F_INVR, -1});
        } else B->aligners.push_back({(int)i, 0, j.flag, -1});
    }
    // pair aligners of identical shape and scoring.  Selection pairs first: INVL+INVR of one -inv job
    // (AGEaligner.cpp:136-154) and mutual -both partners (age_align.cpp:264-278) share a warp, and only the
    // half that the reference would print is enumerated and traced back.
    struct Key { int32_t v[6]; };
    auto key = [&](int a) {
        const age_job &j = jobs[B->aligners[a].job];
        return Key{{j.len1, j.len2, j.match, j.mismatch, j.gap_open, j.gap_extend}};
    };
    auto less = [&](int x, int y) { const Key a = key(x), b = key(y); return std::lexicographical_compare(a.v, a.v + 6, b.v, b.v + 6); };
    auto same = [&](int x, int y) { return memcmp(key(x).v, key(y).v, sizeof(Key)) == 0; };
    auto add_pair = [&](int a0, int a1, int select) {
        HostPair hp{};
        hp.a[0] = a0; hp.a[1] = a1; hp.select = select;
        const age_job &j = jobs[B->aligners[a0].job];
        hp.len1 = j.len1; hp.len2 = j.len2;
        hp.nstrips = (j.len2 + STRIP - 1) / STRIP; hp.P = hp.nstrips * STRIP;
        hp.nb = (j.len1 + W - 1) / W; hp.nsteps = hp.nb + 31; hp.nwin = (hp.nsteps + CKS - 1) / CKS;
        hp.k7 = (B->aligners[a0].flag != F_INDEL) || (a1 >= 0 && B->aligners[a1].flag != F_INDEL);
        const size_t slots = (size_t)hp.nstrips * hp.nsteps;                 // (strip, step) slots of the diagonal-major planes
        // D as 8-bit row offsets when a DP step can add at most 15 (age_device.cuh)
        hp.wide = wide_job[B->aligners[a0].job];
        hp.dfmt = (!hp.wide && !force_raw && std::max((int)j.match, (int)j.mismatch) <= 15) ? DFMT_OFF8 : DFMT_RAW;
        hp.in_bytes = (hp.wide ? 1 : 2) * (align_up((size_t)j.len1, 16) + align_up((size_t)j.len2, 16));   // raw sequences of the halves
        hp.scratch_bytes = align_up(slots * d_planes(hp.dfmt) * 32 * 16, 256) +
                           2 * align_up((size_t)std::max(hp.nstrips - 1, 0) * 3 * (hp.nb + 1) * 16, 256) +
                           2 * align_up(((size_t)hp.P + 2) * 4, 256) + align_up(((size_t)W * hp.nb + 4) * 4, 256) +
                           align_up((size_t)hp.nstrips * hp.nwin * 32 * 4, 256) +
                           2 * align_up((size_t)hp.nstrips * hp.nwin * CKW * 32 * 4, 256);
        hp.scratch_bytes += 2 * align_up(((size_t)hp.nb + 2) * 16, 256) + align_up((size_t)hp.P * 8, 256) + align_up((size_t)hp.P, 256);
        hp.nsets = (select || a1 < 0) ? 1 : 2;
        hp.set_bytes = p2_set_bytes((size_t)hp.nstrips * hp.nwin, (size_t)j.len2);
        hp.scratch_bytes += hp.nsets * hp.set_bytes;
        // Trace pages come out of a pool of the batch: an INDEL aligner looks at a few dozen of its windows (about 5 % of all its
        // pages in the C2 workload), a TDUP / inversion aligner at whole strips of the tie rows (about 10 % in C4).  The pool gets
        // an eighth of the full size or what those strips take; a batch that needs more is run again with a larger pool.
        hp.trace_full = (size_t)hp.nsets * p2_full_granules((size_t)hp.nstrips * hp.nwin, hp.k7);
        // (small matrices touch a larger part of their windows: at least 64 granules per aligner, never more than all pages)
        // (a TDUP / inversion aligner sweeps the whole strips of its tie rows, usually one to three: that many strips' worth of pages)
        const size_t k7_strips = (size_t)3 * hp.nwin * 41 * hp.nsets;
        hp.trace_share = std::min(hp.trace_full, std::max(hp.trace_full / 8, hp.k7 ? k7_strips : (size_t)64 * hp.nsets) + 8);
        hp.cells = (double)j.len1 * j.len2;
        B->pairs.push_back(hp);
    };
    std::vector<int> first_of(n, -1);
    for (size_t a = 0; a < B->aligners.size(); ++a) if (B->aligners[a].role == 0) first_of[B->aligners[a].job] = (int)a;
    std::vector<char> used(B->aligners.size(), 0);
    // wide aligners run one per warp: the selection between partners / INVL and INVR is made on the host (assemble)
    for (size_t a = 0; a < B->aligners.size(); ++a)
        if (wide_job[B->aligners[a].job]) { add_pair((int)a, -1, 0); used[a] = 1; }
    for (size_t a = 0; a < B->aligners.size(); ++a) {
        const Aligner &al = B->aligners[a];
        if (used[a] || al.role != 0) continue;
        const age_job &j = jobs[al.job];
        if (j.flag & F_INV) { add_pair((int)a, (int)a + 1, 1); used[a] = used[a + 1] = 1; continue; }
        const int pj = j.partner;
        if (pj > al.job && (size_t)pj < n && jobs[pj].partner == al.job && B->job_status[pj] == AGE_OK && !wide_job[pj] &&
            !(jobs[pj].flag & F_INV) && same((int)a, first_of[pj])) {
            add_pair((int)a, first_of[pj], 1); used[a] = used[first_of[pj]] = 1;
        }
    }
    std::vector<int> order;
    for (size_t i = 0; i < B->aligners.size(); ++i) if (!used[i]) order.push_back((int)i);
    std::stable_sort(order.begin(), order.end(), less);
    for (size_t i = 0; i < order.size();) {
        if (i + 1 < order.size() && same(order[i + 1], order[i])) { add_pair(order[i], order[i + 1], 0); i += 2; }
        else { add_pair(order[i], -1, 0); i += 1; }
    }
    B->partner_of.assign(n, -1);
    for (size_t i = 0; i < n; ++i) {
        const int pj = jobs[i].partner;
        if (pj >= 0 && (size_t)pj < n && (size_t)pj != i && jobs[pj].partner == (int)i) B->partner_of[i] = pj;
    }
    B->outs.assign(B->aligners.size(), AlignerOut{});
    B->blocks.assign(B->aligners.size(), {});
}

uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// Copy the raw sequences of pair `pi` into the pinned host blob at `h` (device address `d`) and fill its task record.
// The column descriptors and row tables are derived from the raw strings on the device (age_pack_kernel).
void pack_pair(const Batch *B, const age_job *jobs, int pi, char *h, char *d, PairTask &T)
{
    const HostPair &hp = B->pairs[pi];
    memset(&T, 0, sizeof(T));
    T.len1 = hp.len1; T.len2 = hp.len2; T.nstrips = hp.nstrips; T.P = hp.P;
    T.select = hp.select; T.nb = hp.nb; T.nsteps = hp.nsteps; T.nwin = hp.nwin;
    int sc[2][4] = {{1, -1, -1, -1}, {1, -1, -1, -1}};
    size_t off = 0;
    const char *first_seq1 = nullptr;
    for (int k = 0; k < 2; ++k) {
        T.flag[k] = 0; T.out[k] = -1;
        if (hp.a[k] < 0) continue;
        const Aligner &al = B->aligners[hp.a[k]];
        const age_job &j = jobs[al.job];
        T.flag[k] = al.flag; T.out[k] = al.out;
        if (k == 1 && first_seq1 == j.seq1) T.raw1[1] = T.raw1[0];          // -both / -inv halves share the window
        else { memcpy(h + off, j.seq1, hp.len1); T.raw1[k] = d + off; off += align_up(hp.len1, 16); first_seq1 = j.seq1; }
        memcpy(h + off, j.seq2, hp.len2); T.raw2[k] = d + off; off += align_up(hp.len2, 16);
        sc[k][0] = j.match; sc[k][1] = j.mismatch; sc[k][2] = j.gap_open; sc[k][3] = j.gap_extend;
        T.match[k] = j.match; T.mismatch[k] = j.mismatch;
    }
    if (hp.a[1] < 0) { for (int q = 0; q < 4; ++q) sc[1][q] = sc[0][q]; }
    int c0x[2], kabs[2], flip[2], gb[2];
    for (int k = 0; k < 2; ++k) {
        const int go = sc[k][2], ge = sc[k][3];
        const int K = 2 * (ge - go) - 1;
        gb[k] = 2 * go + 1;
        if (K >= 0) { flip[k] = 0; kabs[k] = K; c0x[k] = 2 * go + 1; }
        else        { flip[k] = 1; kabs[k] = -K; c0x[k] = 2 * ge; }
    }
    // kabs multiplies a packed 0/1 word, so both halves share it: build_plan only pairs aligners with
    // identical scoring
    T.c0x = pack2(c0x[0], c0x[1]); T.kabs = (uint32_t)kabs[0]; T.flip = pack2(flip[0], flip[1]); T.gborder = pack2(gb[0], gb[1]);
    T.m2 = pack2(2 * sc[0][0], 2 * sc[1][0]); T.mm2 = pack2(2 * sc[0][1], 2 * sc[1][1]);
    if (hp.wide) {                   // one aligner, plain 32-bit words
        T.c0x = (uint32_t)c0x[0]; T.flip = (uint32_t)flip[0]; T.gborder = (uint32_t)gb[0];
        T.m2 = (uint32_t)(2 * sc[0][0]); T.mm2 = (uint32_t)(2 * sc[0][1]);
    }
    T.dfmt = hp.dfmt; T.wide = hp.wide;
}

void assign_scratch(const HostPair &hp, char *s, PairTask &T)
{
    const size_t slots = (size_t)hp.nstrips * hp.nsteps;
    const size_t bsz = align_up((size_t)std::max(hp.nstrips - 1, 0) * 3 * (hp.nb + 1) * 16, 256);
    const size_t csz = align_up((size_t)hp.nstrips * hp.nwin * CKW * 32 * 4, 256);
    T.D = (uint32_t *)s; s += align_up(slots * d_planes(hp.dfmt) * 32 * 16, 256);
    T.bnd_f = (uint4 *)s; s += bsz;
    T.bnd_r = (uint4 *)s; s += bsz;
    T.Dlast = (uint32_t *)s; s += align_up(((size_t)hp.P + 2) * 4, 256);
    T.Rfirst = (uint32_t *)s; s += align_up(((size_t)hp.P + 2) * 4, 256);
    T.Dtail = (uint32_t *)s; s += align_up(((size_t)W * hp.nb + 4) * 4, 256);
    T.tilemax = (uint32_t *)s; s += align_up((size_t)hp.nstrips * hp.nwin * 32 * 4, 256);
    T.ckpt_f = (uint32_t *)s; s += csz;
    T.ckpt_r = (uint32_t *)s; s += csz;
    const size_t xs = align_up(((size_t)hp.nb + 2) * 16, 256);
    T.xblk_f = (uint4 *)s; s += xs;
    T.xblk_r = (uint4 *)s; s += xs;
    T.ytab = (uint2 *)s; s += align_up((size_t)hp.P * 8, 256);
    T.ycode = (uint8_t *)s; s += align_up((size_t)hp.P, 256);
    T.p2set = s; T.p2stride = hp.set_bytes; T.k7 = hp.k7;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { dev.err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

size_t out_bytes_of(int n_outs, int pool_cap)
{
    return align_up((size_t)n_outs * sizeof(AlignerOut), 256) + align_up((size_t)pool_cap * sizeof(Block), 256) + 1024;
}

bool sync_device(Device &dev)
{
    CK(cudaSetDevice(dev.id));
    for (cudaStream_t st : {dev.st, dev.cp, dev.rb}) if (st) CK(cudaStreamSynchronize(st));
    return true;
}

// stage one sub-batch (pair ids) in a slot of a device: pack, copy inputs (copy stream), lay out scratch
bool stage_on_device(age_ctx *ctx, Batch &B, const age_job *jobs, Device &dev, Slot &sl, const std::vector<int> &pair_ids_in)
{
    CK(cudaSetDevice(dev.id));
    // wide pairs first (they run in their own kernels), then the longest pairs first: the persistent sweep kernel hands out
    // pairs in this order, so its tail is made of short ones
    sl.pair_ids = pair_ids_in;
    std::stable_sort(sl.pair_ids.begin(), sl.pair_ids.end(), [&](int a, int b) {
        if (B.pairs[a].wide != B.pairs[b].wide) return B.pairs[a].wide > B.pairs[b].wide;
        return B.pairs[a].cells * B.pairs[a].nstrips > B.pairs[b].cells * B.pairs[b].nstrips; });
    const std::vector<int> &pair_ids = sl.pair_ids;
    const size_t np = pair_ids.size();
    sl.tasks.assign(np, PairTask{});
    std::vector<size_t> in_off(np + 1, 0);
    sl.sc_off.assign(np + 1, 0);
    int n_outs = 0, ms = 1, ms_wide = 1;
    size_t nwt_total = 0;
    sl.n_wide = 0;
    for (size_t i = 0; i < np; ++i) {
        HostPair &hp = B.pairs[pair_ids[i]];
        sl.n_wide += hp.wide;
        (hp.wide ? ms_wide : ms) = std::max(hp.wide ? ms_wide : ms, hp.nstrips);
        in_off[i + 1] = in_off[i] + hp.in_bytes;
        sl.sc_off[i + 1] = sl.sc_off[i] + hp.scratch_bytes;
        nwt_total += (size_t)hp.nstrips * hp.nwin * hp.nsets;
        for (int k = 0; k < 2; ++k) if (hp.a[k] >= 0) B.aligners[hp.a[k]].out = n_outs++;
    }
    sweep_plan((int)np - sl.n_wide, ms, &sl.nw, &sl.cs);
    sweep_plan(sl.n_wide, ms_wide, &sl.nw_wide, &sl.cs_wide);
    sl.n_outs = n_outs;
    // the trace pool of the sub-batch behind the scratch of its pairs
    size_t share = 0;
    sl.tpool_full = P2_DUMP_GRANULES;
    for (size_t i = 0; i < np; ++i) { share += B.pairs[pair_ids[i]].trace_share; sl.tpool_full += B.pairs[pair_ids[i]].trace_full; }
    size_t gran = std::min(sl.tpool_full, std::max<size_t>((size_t)((double)share * dev.trace_scale) + P2_DUMP_GRANULES, (size_t)16384));   // at least 64 MB
    if (const char *e = getenv("AGE_TRACE_POOL_GRANULES")) gran = std::max<size_t>(P2_DUMP_GRANULES + 64, (size_t)atol(e));   // test knob: a small pool forces the overflow re-run
    sl.tpool_granules = (uint32_t)std::min<size_t>(gran, 0xFFFFFFF0u);
    sl.tpool_off = align_up(sl.sc_off[np], P2_GRANULE);
    sl.sc_bytes = sl.tpool_off + (size_t)sl.tpool_granules * P2_GRANULE;
    sl.queue_cap = (int)std::min<size_t>(2 * nwt_total + 64, (size_t)1 << 24);
    if (const char *e = getenv("AGE_QUEUE_CAP")) sl.queue_cap = std::max(1, atoi(e));   // test knob: a short work list forces on-demand re-sweeps
    const size_t task_bytes = align_up(np * sizeof(PairTask), 256);
    sl.in_bytes = task_bytes + in_off[np];
    if (!sl.hin.ensure(sl.in_bytes)) { dev.err = "pinned host allocation failed"; return false; }
    if (!sl.in.ensure(sl.in_bytes)) { dev.err = "device allocation failed (inputs)"; return false; }
    if (sl.sc_bytes + 256 > dev.scratch.cap) {
        CK(cudaStreamSynchronize(dev.st));               // the batch in flight still works in the arena that is about to move
        if (!dev.scratch.ensure(sl.sc_bytes + 256)) { dev.err = "device allocation failed (scratch)"; return false; }
    }
    sl.pool_cap = n_outs * 96 + 4096;
    if (const char *e = getenv("AGE_POOL_CAP")) sl.pool_cap = std::max(1, atoi(e));     // test knob: a small block pool forces the overflow re-run
    sl.out_bytes = out_bytes_of(n_outs, sl.pool_cap);
    if (!sl.outbuf.ensure(sl.out_bytes)) { dev.err = "device allocation failed (outputs)"; return false; }
    if (!sl.hout.ensure(sl.out_bytes)) { dev.err = "pinned host allocation failed"; return false; }
    if (!sl.aux.ensure(align_up(np * sizeof(PairHdr), 256) + (size_t)sl.queue_cap * sizeof(uint2))) { dev.err = "device allocation failed (work list)"; return false; }
    parallel_for(ctx->host_threads, np, [&](size_t i) {
        pack_pair(&B, jobs, pair_ids[i], sl.hin.ptr + task_bytes + in_off[i], sl.in.ptr + task_bytes + in_off[i], sl.tasks[i]);
        assign_scratch(B.pairs[pair_ids[i]], dev.scratch.ptr + sl.sc_off[i], sl.tasks[i]);
    });
    memcpy(sl.hin.ptr, sl.tasks.data(), np * sizeof(PairTask));
    CK(cudaEventRecord(sl.ev[0], dev.cp));
    CK(cudaMemcpyAsync(sl.in.ptr, sl.hin.ptr, sl.in_bytes, cudaMemcpyHostToDevice, dev.cp));
    CK(cudaEventRecord(sl.ev[1], dev.cp));
    CK(cudaStreamWaitEvent(dev.st, sl.ev[1], 0));
    launch_pack((const PairTask *)sl.in.ptr, (int)np, dev.st);
    return true;
}

struct OutLayout { AlignerOut *outs; Block *pool; int32_t *counters; };
OutLayout out_layout(const Slot &sl, char *base)
{
    OutLayout L;
    L.outs = (AlignerOut *)base;
    L.pool = (Block *)(base + align_up((size_t)sl.n_outs * sizeof(AlignerOut), 256));
    L.counters = (int32_t *)((char *)L.pool + align_up((size_t)sl.pool_cap * sizeof(Block), 256));   // 32 ints
    return L;
}

// all kernels of the slot's sub-batch on the compute stream
bool run_on_device(Device &dev, Slot &sl)
{
    CK(cudaSetDevice(dev.id));
    OutLayout L = out_layout(sl, sl.outbuf.ptr);
    CK(cudaEventRecord(sl.ev[2], dev.st));
    BatchLaunch b{};
    b.pairs = (const PairTask *)sl.in.ptr; b.n_pairs = (int)sl.tasks.size(); b.n_wide = sl.n_wide; b.nw = sl.nw; b.nw_wide = sl.nw_wide; b.cs = sl.cs; b.cs_wide = sl.cs_wide;
    b.outs = L.outs; b.pool = L.pool; b.pool_cap = sl.pool_cap; b.counters = L.counters;
    b.hdr = (PairHdr *)sl.aux.ptr;
    b.queue = (uint2 *)(sl.aux.ptr + align_up(sl.tasks.size() * sizeof(PairHdr), 256)); b.queue_cap = sl.queue_cap;
    b.tpool = dev.scratch.ptr + sl.tpool_off; b.tpool_granules = sl.tpool_granules;
    launch_sweeps(b, dev.st);
    CK(cudaEventRecord(sl.ev[3], dev.st));
    launch_pass2(b, dev.st);
    CK(cudaEventRecord(sl.ev[4], dev.st));
    sl.launches = launches_per_batch(b);
    CK(cudaGetLastError());
    return true;
}

// result read-back on the copy stream, behind the kernels
bool enqueue_fetch(Device &dev, Slot &sl)
{
    CK(cudaSetDevice(dev.id));
    CK(cudaStreamWaitEvent(dev.rb, sl.ev[4], 0));
    CK(cudaEventRecord(sl.ev[5], dev.rb));
    CK(cudaMemcpyAsync(sl.hout.ptr, sl.outbuf.ptr, sl.out_bytes, cudaMemcpyDeviceToHost, dev.rb));
    CK(cudaEventRecord(sl.ev[6], dev.rb));
    return true;
}

// The block pool or the trace pool of the slot was too small: run its sub-batch again with larger ones.  A later batch may
// have used (or moved) the scratch arena in the meantime, so the sweeps are repeated as well; the packed inputs are still resident.
bool rerun_with_pools(Batch &B, Device &dev, Slot &sl, int pool_cap, size_t tpool_granules)
{
    if (!sync_device(dev)) return false;
    sl.pool_cap = pool_cap;
    sl.out_bytes = out_bytes_of(sl.n_outs, sl.pool_cap);
    if (!sl.outbuf.ensure(sl.out_bytes) || !sl.hout.ensure(sl.out_bytes)) { dev.err = "allocation failed (block pool)"; return false; }
    sl.tpool_granules = (uint32_t)std::min<size_t>(std::min(tpool_granules, sl.tpool_full), 0xFFFFFFF0u);
    sl.sc_bytes = sl.tpool_off + (size_t)sl.tpool_granules * P2_GRANULE;
    if (sl.sc_bytes + 256 > dev.scratch.cap && !dev.scratch.ensure(sl.sc_bytes + 256)) { dev.err = "device allocation failed (scratch)"; return false; }
    for (size_t i = 0; i < sl.tasks.size(); ++i) assign_scratch(B.pairs[sl.pair_ids[i]], dev.scratch.ptr + sl.sc_off[i], sl.tasks[i]);
    memcpy(sl.hin.ptr, sl.tasks.data(), sl.tasks.size() * sizeof(PairTask));
    CK(cudaMemcpyAsync(sl.in.ptr, sl.hin.ptr, sl.tasks.size() * sizeof(PairTask), cudaMemcpyHostToDevice, dev.st));
    launch_pack((const PairTask *)sl.in.ptr, (int)sl.tasks.size(), dev.st);
    return run_on_device(dev, sl) && enqueue_fetch(dev, sl);
}

// wait for the slot's results and gather them into the batch
bool finish_fetch(Batch &B, Device &dev, Slot &sl, age_stats &st, int host_threads = 1)
{
    CK(cudaSetDevice(dev.id));
    for (int attempt = 0; attempt < 8; ++attempt) {
        CK(cudaEventSynchronize(sl.ev[6]));
        OutLayout H = out_layout(sl, sl.hout.ptr);
        const size_t tneed = (size_t)(uint32_t)H.counters[3] + P2_DUMP_GRANULES;
        if (H.counters[1] > sl.pool_cap || tneed > sl.tpool_granules) {       // a pool was too small: run again with larger ones
            const int pc = H.counters[1] > sl.pool_cap ? H.counters[1] + 4096 : sl.pool_cap;
            const size_t tg = tneed > sl.tpool_granules ? std::max(2 * (size_t)sl.tpool_granules, tneed + tneed / 4) : sl.tpool_granules;
            if (tneed > sl.tpool_granules) dev.trace_scale = std::min(8.0, dev.trace_scale * std::max(1.5, 1.25 * (double)tneed / (double)sl.tpool_granules));
            if (g_hprof) fprintf(stderr, "[host prof] pool overflow: blocks %d of %d, trace granules %zu of %u -> run again\n", H.counters[1], sl.pool_cap, tneed, sl.tpool_granules);
            if (!rerun_with_pools(B, dev, sl, pc, tg)) return false;
            st.launches += 1 + sl.launches; st.fill_launches += 1; st.pool_reruns += 1;
            continue;
        }
        parallel_for(host_threads, sl.pair_ids.size(), [&](size_t i) {     // pairs are disjoint: gathered on all host cores
            const HostPair &hp = B.pairs[sl.pair_ids[i]];
            for (int k = 0; k < 2; ++k) {
                if (hp.a[k] < 0) continue;
                const Aligner &al = B.aligners[hp.a[k]];
                const AlignerOut &o = H.outs[al.out];
                memcpy(&B.outs[hp.a[k]], &o, offsetof(AlignerOut, bp) + sizeof(o.bp[0]) * (size_t)std::max(o.n_bp, 0));
                std::vector<age_block> &b = B.blocks[hp.a[k]];
                b.clear();
                for (int q = 0; q < 4; ++q)
                    for (int x = 0; x < o.n_blk[q]; ++x) {
                        const Block &s = H.pool[o.blk_off[q] + x];
                        b.push_back(age_block{s.start1, s.start2, s.length});
                    }
            }
        });
        float ms = 0;
        st.h2d_bytes += sl.in_bytes; st.d2h_bytes += sl.out_bytes;
        if (cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[1]) == cudaSuccess) st.h2d_ms += ms;
        if (cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[3]) == cudaSuccess) st.fill_ms += ms;
        if (cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[4]) == cudaSuccess) st.kernel_ms += ms;
        if (cudaEventElapsedTime(&ms, sl.ev[5], sl.ev[6]) == cudaSuccess) st.d2h_ms += ms;
        st.launches += 1 + sl.launches; st.fill_launches += 1;          // pack + the batch's kernels
        return true;
    }
    dev.err = "block / trace pool overflow";
    return false;
}

// Longest-processing-time-first deal of the pairs of a plan across nd devices (cost = DP area; a wide pair costs about
// as much as a packed one per area: one aligner at ~2.2x the instructions instead of two).  Each device keeps plan order.
std::vector<std::vector<int>> deal_pairs(const Batch &B, int nd)
{
    std::vector<int> order(B.pairs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::vector<std::vector<int>> per_dev(nd);
    if (nd == 1) { per_dev[0] = order; return per_dev; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return B.pairs[a].cells > B.pairs[b].cells; });
    std::vector<double> load(nd, 0.0);
    for (int pi : order) {
        int best = 0;
        for (int d = 1; d < nd; ++d) if (load[d] < load[best]) best = d;
        per_dev[best].push_back(pi); load[best] += B.pairs[pi].cells;
    }
    for (auto &v : per_dev) std::sort(v.begin(), v.end());
    return per_dev;
}

// split the pairs of the plan into per-device sub-batches that fit the scratch budget
std::vector<std::vector<std::vector<int>>> make_subbatches(age_ctx *ctx, const Batch &B, int slot)
{
    const int nd = (int)ctx->devs.size();
    std::vector<std::vector<std::vector<int>>> res(nd);
    std::vector<size_t> budget(nd, 0);
    auto query_budget = [&](int d) {             // cudaMemGetInfo costs milliseconds: only asked when a batch outgrows the arenas
        Device &dev = ctx->devs[d];
        cudaSetDevice(dev.id);
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        fr += dev.scratch.cap + dev.slot[slot].in.cap + dev.slot[slot].outbuf.cap + dev.slot[slot].aux.cap;
        budget[d] = (size_t)(0.85 * (double)fr);
    };
    std::vector<std::vector<int>> per_dev = deal_pairs(B, nd);
    const char *cap_env = getenv("AGE_MEM_BUDGET_MB");               // test knob: a small budget forces several sub-batches
    for (int d = 0; d < nd; ++d) {
        size_t sc = 256, in = align_up(per_dev[d].size() * sizeof(PairTask), 256);
        size_t share = P2_DUMP_GRANULES;
        for (int pi : per_dev[d]) { sc += B.pairs[pi].scratch_bytes; in += B.pairs[pi].in_bytes; share += B.pairs[pi].trace_share; }
        sc += std::max<size_t>(share, 16384) * P2_GRANULE + P2_GRANULE;
        if (!cap_env && sc <= ctx->devs[d].scratch.cap && in <= ctx->devs[d].slot[slot].in.cap) {        // fits what is already allocated
            if (!per_dev[d].empty()) res[d].push_back(per_dev[d]);
            continue;
        }
        query_budget(d);
        if (cap_env) budget[d] = (size_t)std::max(1L, atol(cap_env)) << 20;
        std::vector<int> cur; size_t used = 0;
        for (int pi : per_dev[d]) {
            const size_t need = B.pairs[pi].scratch_bytes + B.pairs[pi].in_bytes + 4096 + B.pairs[pi].trace_share * P2_GRANULE;
            if (!cur.empty() && used + need > budget[d]) { res[d].push_back(cur); cur.clear(); used = 0; }
            cur.push_back(pi); used += need;
        }
        if (!cur.empty()) res[d].push_back(cur);
    }
    return res;
}

void account(Batch &B)
{
    B.stats.n_aligners = B.aligners.size();
    B.stats.cell_updates = 0; B.stats.sweep_bytes = 0;
    for (const HostPair &hp : B.pairs) {
        for (int k = 0; k < 2; ++k) if (hp.a[k] >= 0) B.stats.cell_updates += (uint64_t)(2.0 * hp.cells);
        // D written by the forward and read by the reverse sweep (lanes whose rows lie beyond the contig skip it), a
        // checkpoint per window and direction, boundary rows written and read per direction, one tile maximum per lane and window
        const double live = std::min(1.0, (hp.len2 / 8 + 1) / (32.0 * hp.nstrips));
        B.stats.sweep_bytes += (uint64_t)(2.0 * live * hp.nstrips * hp.nsteps * d_planes(hp.dfmt) * 512.0) +
                               2ull * hp.nstrips * hp.nwin * CKW * 128 + 4ull * std::max(hp.nstrips - 1, 0) * 3 * (hp.nb + 1) * 16 +
                               (uint64_t)hp.nstrips * hp.nwin * 128;
    }
}

void assemble(const Batch &B, age_result *out, size_t n, int host_threads = 1)
{
    std::vector<int> main_of(n, -1), aux_of(n, -1);
    for (size_t a = 0; a < B.aligners.size(); ++a) {
        const Aligner &al = B.aligners[a];
        (al.role ? aux_of : main_of)[al.job] = (int)a;
    }
    parallel_for(host_threads, n, [&](size_t i) {
        age_result &r = out[i];
        memset(&r, 0, offsetof(age_result, bp));          // header only: bp rows beyond n_bp are not part of the result
        memset(&r.counts, 0, sizeof(r.counts));
        r.alt_index = -1; r.n_left = r.n_right = r.n_alt_left = r.n_alt_right = 0;
        r.left = r.right = r.alt_left = r.alt_right = nullptr;
        r.status = B.job_status[i];
        r.aux_score = -1;
        if (r.status != AGE_OK) return;
        const int m = main_of[i], x = aux_of[i];
        const AlignerOut &om = B.outs[m];
        r.main_score = om.score;
        int pick = m;
        if (x >= 0) {
            r.aux_score = B.outs[x].score;
            if (B.outs[x].score > om.score) { pick = x; r.used_aux = 1; }       // AGEaligner.cpp:151,181-184
        }
        const AlignerOut &o = B.outs[pick];
        if (o.status != 0) { r.status = AGE_ERR_CUDA; return; }
        r.score = o.score; r.flag = B.aligners[pick].flag;
        if (o.n_bp < 0) return;                                    // lost the -both selection: score only
        r.detail = 1;
        r.n_bp = o.n_bp;
        memcpy(r.bp, o.bp, sizeof(r.bp[0]) * (size_t)std::max(o.n_bp, 0));
        r.alt_index = o.alt_index;
        r.n_left = o.n_blk[0]; r.n_right = o.n_blk[1]; r.n_alt_left = o.n_blk[2]; r.n_alt_right = o.n_blk[3];
        const age_block *b = B.blocks[pick].data();
        r.left = b; r.right = b + r.n_left; r.alt_left = r.right + r.n_right; r.alt_right = r.alt_left + r.n_alt_left;
        age_counts &c = r.counts;
        for (int k = 0; k < 2; ++k) { c.n_aligned[k] = o.call[CALL_ALIGNED + k]; c.n_identic[k] = o.call[CALL_IDENTIC + k]; c.n_gap[k] = o.call[CALL_GAP + k]; }
        for (int k = 0; k < 3; ++k) { c.homo_at1[k] = o.call[CALL_AT1 + k]; c.homo_at2[k] = o.call[CALL_AT2 + k]; }
        c.homo_out1 = o.call[CALL_OUT1]; c.homo_out2 = o.call[CALL_OUT2]; c.homo_in1 = o.call[CALL_IN1]; c.homo_in2 = o.call[CALL_IN2];
    });
    // Mutual -both partners that did not share a warp (wide jobs) both come back complete: keep the details of the one
    // age_align prints only (the lower index unless the other scores strictly higher, age_align.cpp:273), as the
    // on-device selection does.
    for (size_t i = 0; i < n; ++i) {
        const int pj = B.partner_of[i];
        if (pj <= (int)i) continue;
        age_result &a = out[i], &b = out[pj];
        if (a.status != AGE_OK || b.status != AGE_OK || !a.detail || !b.detail) continue;
        age_result &lose = b.score > a.score ? a : b;
        lose.detail = 0; lose.n_bp = 0; lose.alt_index = -1; lose.used_aux = 0;
        lose.n_left = lose.n_right = lose.n_alt_left = lose.n_alt_right = 0;
        lose.left = lose.right = lose.alt_left = lose.alt_right = nullptr;
        memset(&lose.counts, 0, sizeof(lose.counts));
    }
}

// devices run concurrently: report the slowest one for the time-like fields, sums for the counters
void merge_stats(age_stats &to, const std::vector<age_stats> &per_dev)
{
    for (const age_stats &s : per_dev) {
        to.kernel_ms = std::max(to.kernel_ms, s.kernel_ms); to.fill_ms = std::max(to.fill_ms, s.fill_ms);
        to.h2d_ms = std::max(to.h2d_ms, s.h2d_ms); to.d2h_ms = std::max(to.d2h_ms, s.d2h_ms);
        to.h2d_bytes += s.h2d_bytes; to.d2h_bytes += s.d2h_bytes; to.launches += s.launches; to.fill_launches += s.fill_launches; to.pool_reruns += s.pool_reruns;
    }
}

template <class F> bool for_each_device(age_ctx *ctx, F f)        // f(d) -> bool, one host thread per device
{
    const int nd = (int)ctx->devs.size();
    std::vector<char> ok(nd, 1);
    if (nd == 1) ok[0] = f(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; ++d) th.emplace_back([&, d] { ok[d] = f(d); });
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < nd; ++d) if (!ok[d]) { ctx->err = ctx->devs[d].err; return false; }
    return true;
}

// plan + stage + launch; the batch lands in slot ctx->head
int submit(age_ctx *ctx, const age_job *jobs, size_t n)
{
    const int s = ctx->head;
    Batch &B = ctx->batch[s];
    if (B.state != 0) { ctx->err = "two batches are already in flight: call age_collect first"; return AGE_ERR_ARG; }
    const double t0 = wall_ms();
    B.stats = age_stats{};
    build_plan(&B, jobs, n);
    const double t1 = wall_ms();
    auto sub = make_subbatches(ctx, B, s);
    const int nd = (int)ctx->devs.size();
    bool single = true;
    for (int d = 0; d < nd; ++d) single = single && sub[d].size() <= 1;
    B.on_dev.assign(nd, 0);
    std::vector<age_stats> dstat(nd);
    bool ok;
    if (single) {                 // the common case: asynchronous, collected later
        ok = for_each_device(ctx, [&](int d) {
            if (sub[d].empty()) return true;
            Device &dev = ctx->devs[d];
            B.on_dev[d] = 1;
            return stage_on_device(ctx, B, jobs, dev, dev.slot[s], sub[d][0]) && run_on_device(dev, dev.slot[s]) && enqueue_fetch(dev, dev.slot[s]);
        });
        B.state = 1;
    } else {                      // larger than the device memory: piece by piece, synchronously
        ok = for_each_device(ctx, [&](int d) {
            Device &dev = ctx->devs[d];
            for (auto &ids : sub[d])
                if (!stage_on_device(ctx, B, jobs, dev, dev.slot[s], ids) || !run_on_device(dev, dev.slot[s]) || !enqueue_fetch(dev, dev.slot[s]) ||
                    !finish_fetch(B, dev, dev.slot[s], dstat[d])) return false;
            return true;
        });
        merge_stats(B.stats, dstat);
        B.state = 2;
    }
    if (!ok) {                    // other devices may already have copies or kernels of this batch in flight: drain them before the slot is reused
        for (Device &dev : ctx->devs) sync_device(dev);
        B.state = 0;
        return AGE_ERR_CUDA;
    }
    ctx->head ^= 1; ctx->in_flight += 1;
    if (g_hprof) fprintf(stderr, "[host prof] submit: plan %.2f ms, sub-batching + stage + launch %.2f ms\n", t1 - t0, wall_ms() - t1);
    return 0;
}

// wait for the oldest batch in flight and hand its results out
int collect(age_ctx *ctx, age_result *out, size_t n)
{
    if (ctx->in_flight <= 0) { ctx->err = "no batch in flight"; return AGE_ERR_ARG; }
    const int s = ctx->in_flight == 2 ? ctx->head : ctx->head ^ 1;
    Batch &B = ctx->batch[s];
    if (n != B.n_jobs) { ctx->err = "age_collect: n differs from the submitted batch"; return AGE_ERR_ARG; }
    const double t0 = wall_ms();
    bool ok = true;
    if (B.state == 1) {
        const int nd = (int)ctx->devs.size();
        std::vector<age_stats> dstat(nd);
        ok = for_each_device(ctx, [&](int d) { return !B.on_dev[d] || finish_fetch(B, ctx->devs[d], ctx->devs[d].slot[s], dstat[d], std::max(1, ctx->host_threads / (int)ctx->devs.size())); });
        merge_stats(B.stats, dstat);
    }
    B.state = 0; ctx->in_flight -= 1;
    if (!ok) { for (Device &dev : ctx->devs) sync_device(dev); return AGE_ERR_CUDA; }
    const double t1 = wall_ms();
    account(B);
    assemble(B, out, n, ctx->host_threads);
    ctx->stats = B.stats;
    if (g_hprof) fprintf(stderr, "[host prof] collect: wait + gather %.2f ms, assemble %.2f ms; kernels %.2f ms\n", t1 - t0, wall_ms() - t1, B.stats.kernel_ms);
    return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" age_ctx *age_create(const int *device_ids, int n_dev)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        g_create_error = "no usable CUDA device (the AGE realigner has no CPU fallback)";
        return nullptr;
    }
    if (n_dev <= 0 || !device_ids) n_dev = 1;
    const double t_create = wall_ms();
    age_ctx *ctx = new age_ctx();
    ctx->host_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    for (int i = 0; i < n_dev; ++i) {
        Device d;
        d.id = device_ids ? device_ids[i] : 0;
        if (d.id < 0 || d.id >= count) { g_create_error = "invalid device id"; delete ctx; return nullptr; }
        ctx->devs.push_back(d);
    }
    // one host thread per device: creating a context takes a few hundred milliseconds, eight of them one after the other seconds
    const bool all_ok = for_each_device(ctx, [&](int di) {
        Device &d = ctx->devs[di];
        bool ok = cudaSetDevice(d.id) == cudaSuccess && cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&d.cp, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&d.rb, cudaStreamNonBlocking) == cudaSuccess;
        for (Slot &sl : d.slot)
            for (int e = 0; ok && e < 7; ++e) ok = cudaEventCreate(&sl.ev[e]) == cudaSuccess;
        if (!ok) d.err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
        return ok;
    });
    if (!all_ok) { g_create_error = ctx->err; age_destroy(ctx); return nullptr; }
    if (g_hprof) fprintf(stderr, "[host prof] age_create %.1f ms\n", wall_ms() - t_create);
    return ctx;
}

extern "C" void age_destroy(age_ctx *ctx)
{
    if (!ctx) return;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        for (cudaStream_t st : {d.st, d.cp, d.rb}) if (st) cudaStreamSynchronize(st);
        d.scratch.release();
        for (Slot &sl : d.slot) {
            sl.in.release(); sl.outbuf.release(); sl.aux.release(); sl.hin.release(); sl.hout.release();
            for (auto &e : sl.ev) if (e) cudaEventDestroy(e);
        }
        for (cudaStream_t st : {d.st, d.cp, d.rb}) if (st) cudaStreamDestroy(st);
    }
    delete ctx;
}

extern "C" int age_reserve(age_ctx *ctx, size_t scratch_bytes)
{
    if (!ctx) return AGE_ERR_ARG;
    if (ctx->in_flight) { ctx->err = "age_reserve: batches are in flight"; return AGE_ERR_ARG; }
    const bool ok = for_each_device(ctx, [&](int di) {
        Device &dev = ctx->devs[di];
        if (cudaSetDevice(dev.id) != cudaSuccess) { dev.err = "cudaSetDevice failed"; return false; }
        size_t want = scratch_bytes;
        if (!want) { size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); want = (size_t)(0.8 * (double)(fr + dev.scratch.cap)); }
        if (want > dev.scratch.cap) {
            cudaStreamSynchronize(dev.st);
            dev.scratch.release();
            if (cudaMalloc(&dev.scratch.ptr, want) != cudaSuccess) { cudaGetLastError(); dev.err = "age_reserve: device allocation failed"; return false; }
            dev.scratch.cap = want;
        }
        if (!scratch_bytes)            // "all of it": the staging buffers of both slots too, sized for a batch of a few thousand candidates
            for (Slot &sl : dev.slot) {
                const bool fine = sl.hin.ensure((size_t)128 << 20) && sl.in.ensure((size_t)128 << 20) && sl.hout.ensure((size_t)64 << 20) &&
                                  sl.outbuf.ensure((size_t)64 << 20) && sl.aux.ensure((size_t)16 << 20);
                if (!fine) { dev.err = "age_reserve: staging buffer allocation failed"; return false; }
            }
        return true;
    });
    return ok ? 0 : AGE_ERR_CUDA;
}

extern "C" const char *age_last_error(const age_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int age_submit(age_ctx *ctx, const age_job *jobs, size_t n)
{
    if (!ctx || (!jobs && n)) return AGE_ERR_ARG;
    ctx->staged = false;
    return submit(ctx, jobs, n);
}

extern "C" int age_collect(age_ctx *ctx, age_result *out, size_t n)
{
    if (!ctx || (!out && n)) return AGE_ERR_ARG;
    return collect(ctx, out, n);
}

extern "C" int age_in_flight(const age_ctx *ctx) { return ctx ? ctx->in_flight : 0; }

extern "C" int age_align_batch(age_ctx *ctx, const age_job *jobs, size_t n, age_result *out)
{
    if (!ctx || (!jobs && n) || (!out && n)) return AGE_ERR_ARG;
    if (ctx->in_flight) { ctx->err = "age_align_batch: batches are in flight (age_submit without age_collect)"; return AGE_ERR_ARG; }
    ctx->staged = false;
    const double tb = wall_ms();
    int rc = submit(ctx, jobs, n);
    if (rc == 0) rc = collect(ctx, out, n);
    if (g_hprof) fprintf(stderr, "[host prof] age_align_batch total %.2f ms\n", wall_ms() - tb);
    return rc;
}

// ---- split form: inputs staged once, kernels repeatable (measurement) --------------------------------------
extern "C" int age_stage(age_ctx *ctx, const age_job *jobs, size_t n)
{
    if (!ctx || (!jobs && n)) return AGE_ERR_ARG;
    if (ctx->in_flight) { ctx->err = "age_stage: batches are in flight"; return AGE_ERR_ARG; }
    ctx->staged = false;
    ctx->head = 0;
    Batch &B = ctx->batch[0];
    B.stats = age_stats{};
    build_plan(&B, jobs, n);
    auto sub = make_subbatches(ctx, B, 0);
    B.on_dev.assign(ctx->devs.size(), 0);
    for (size_t d = 0; d < ctx->devs.size(); ++d) {
        if (sub[d].size() > 1) { ctx->err = "batch does not fit in device memory in one piece; use age_align_batch"; return AGE_ERR_ARG; }
        Device &dev = ctx->devs[d];
        Slot &sl = dev.slot[0];
        if (sub[d].empty()) { sl.pair_ids.clear(); sl.tasks.clear(); sl.n_outs = 0; continue; }
        if (!stage_on_device(ctx, B, jobs, dev, sl, sub[d][0])) { ctx->err = dev.err; return AGE_ERR_CUDA; }
        B.on_dev[d] = 1;
        B.stats.h2d_bytes += sl.in_bytes;
    }
    for (Device &dev : ctx->devs) { cudaSetDevice(dev.id); cudaStreamSynchronize(dev.cp); cudaStreamSynchronize(dev.st); }
    ctx->stats = B.stats;
    ctx->staged = true;
    return 0;
}

extern "C" int age_run_staged(age_ctx *ctx)
{
    if (!ctx || !ctx->staged) return AGE_ERR_ARG;
    Batch &B = ctx->batch[0];
    for (size_t d = 0; d < ctx->devs.size(); ++d) {
        Device &dev = ctx->devs[d];
        if (!B.on_dev[d]) continue;
        if (!run_on_device(dev, dev.slot[0])) { ctx->err = dev.err; return AGE_ERR_CUDA; }
    }
    ctx->stats.kernel_ms = 0; ctx->stats.fill_ms = 0;
    for (size_t d = 0; d < ctx->devs.size(); ++d) {
        Device &dev = ctx->devs[d];
        if (!B.on_dev[d]) continue;
        Slot &sl = dev.slot[0];
        cudaSetDevice(dev.id);
        if (cudaEventSynchronize(sl.ev[4]) != cudaSuccess) { ctx->err = cudaGetErrorString(cudaGetLastError()); return AGE_ERR_CUDA; }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[4]) == cudaSuccess) ctx->stats.kernel_ms = std::max(ctx->stats.kernel_ms, (double)ms);
        if (cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[3]) == cudaSuccess) ctx->stats.fill_ms = std::max(ctx->stats.fill_ms, (double)ms);
    }
    for (size_t d = 0; d < ctx->devs.size(); ++d) if (B.on_dev[d]) ctx->stats.launches += ctx->devs[d].slot[0].launches;   // packed at stage time
    ctx->stats.fill_launches += 1;
    return 0;
}

extern "C" int age_run_staged_many(age_ctx *ctx, int reps, double *total_ms)
{
    if (!ctx || !ctx->staged || reps < 1) return AGE_ERR_ARG;
    Batch &B = ctx->batch[0];
    const size_t nd = ctx->devs.size();
    std::vector<cudaEvent_t> e0(nd, nullptr);
    for (size_t d = 0; d < nd; ++d) {
        Device &dev = ctx->devs[d];
        if (!B.on_dev[d]) continue;
        cudaSetDevice(dev.id);
        cudaEventCreate(&e0[d]);
        cudaEventRecord(e0[d], dev.st);
    }
    int rc = 0;
    for (int r = 0; r < reps && rc == 0; ++r)            // repeats are queued back to back on the compute stream
        for (size_t d = 0; d < nd; ++d)
            if (B.on_dev[d] && !run_on_device(ctx->devs[d], ctx->devs[d].slot[0])) { ctx->err = ctx->devs[d].err; rc = AGE_ERR_CUDA; break; }
    double worst = 0;
    for (size_t d = 0; d < nd; ++d) {
        Device &dev = ctx->devs[d];
        if (!B.on_dev[d]) continue;
        cudaSetDevice(dev.id);
        float ms = 0;
        if (cudaEventSynchronize(dev.slot[0].ev[4]) != cudaSuccess) { ctx->err = cudaGetErrorString(cudaGetLastError()); rc = AGE_ERR_CUDA; }
        else if (cudaEventElapsedTime(&ms, e0[d], dev.slot[0].ev[4]) == cudaSuccess) worst = std::max(worst, (double)ms);
        cudaEventDestroy(e0[d]);
        if (rc == 0) ctx->stats.launches += (uint64_t)reps * dev.slot[0].launches;
    }
    if (total_ms) *total_ms = worst;
    ctx->stats.kernel_ms = worst / reps;
    return rc;
}

extern "C" int age_fetch(age_ctx *ctx, age_result *out, size_t n)
{
    if (!ctx || !ctx->staged || n != ctx->batch[0].n_jobs || (!out && n)) return AGE_ERR_ARG;
    Batch &B = ctx->batch[0];
    age_stats scratch_stats{};
    for (size_t d = 0; d < ctx->devs.size(); ++d) {
        Device &dev = ctx->devs[d];
        if (!B.on_dev[d]) continue;
        if (!enqueue_fetch(dev, dev.slot[0]) || !finish_fetch(B, dev, dev.slot[0], scratch_stats)) { ctx->err = dev.err; return AGE_ERR_CUDA; }
    }
    ctx->stats.d2h_bytes = scratch_stats.d2h_bytes;
    account(B);
    ctx->stats.n_aligners = B.stats.n_aligners; ctx->stats.cell_updates = B.stats.cell_updates; ctx->stats.sweep_bytes = B.stats.sweep_bytes;
    assemble(B, out, n);
    return 0;
}

extern "C" int age_plan(const age_job *jobs, size_t n, int n_dev, int32_t *device_of_job, int32_t *arith_of_job)
{
    if ((!jobs && n) || n_dev < 1 || !device_of_job) return AGE_ERR_ARG;
    Batch B;
    build_plan(&B, jobs, n);
    for (size_t i = 0; i < n; ++i) { device_of_job[i] = -1; if (arith_of_job) arith_of_job[i] = B.job_status[i] == AGE_OK ? 16 : B.job_status[i]; }
    const auto per_dev = deal_pairs(B, n_dev);
    for (int d = 0; d < n_dev; ++d)
        for (int pi : per_dev[d])
            for (int k = 0; k < 2; ++k) {
                if (B.pairs[pi].a[k] < 0) continue;
                const Aligner &al = B.aligners[B.pairs[pi].a[k]];
                if (al.role == 0) device_of_job[al.job] = d;
                if (arith_of_job && B.pairs[pi].wide) arith_of_job[al.job] = 32;
            }
    return 0;
}

extern "C" int age_get_stats(const age_ctx *ctx, age_stats *st)
{
    if (!ctx || !st) return AGE_ERR_ARG;
    *st = ctx->stats;
    return 0;
}

extern "C" const char *age_version(void) { return "age-b200 0.2 (sm_100a)"; }
