// age_engine.cu -- host side of the C ABI declared in include/age_b200.h.
//
// Replaces the serial candidate loop of the reference (age_align.cpp:231-293, VarUnit.cpp:271-299) with
// batched GPU execution: jobs are validated, expanded into single-mode aligners (-inv = INVL + auxiliary
// INVR, AGEaligner.cpp:72-77,136-144), paired by shape, packed, swept and enumerated on the device(s), and
// the compact results (score, breakpoints, diagonal blocks) are gathered on the host in input order.
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
    int nsets; size_t set_bytes;  // pass-2 trace plane sets (one per aligner that gets breakpoints)
    size_t in_bytes, scratch_bytes;
    double cells;
};

struct Device {
    int id = 0;
    cudaStream_t st = nullptr;
    Arena in, scratch, outbuf, aux;
    int queue_cap = 0;
    PinnedBuf hin, hout;
    cudaEvent_t ev[6] = {};
    // the sub-batch currently resident on this device
    std::vector<int> pair_ids;
    std::vector<PairTask> tasks;
    int n_outs = 0, pool_cap = 0;
    size_t in_bytes = 0;
    std::string err;
};

} // namespace

struct age_ctx {
    std::vector<Device> devs;
    std::string err;
    age_stats stats{};
    // plan of the current batch
    const age_job *jobs = nullptr; size_t n_jobs = 0;
    std::vector<int> job_status;
    std::vector<Aligner> aligners;
    std::vector<HostPair> pairs;
    std::vector<AlignerOut> outs;                 // per aligner (global index)
    std::vector<std::vector<age_block>> blocks;   // per aligner: left | right | alt_left | alt_right
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

// 16-bit range of the packed (2*score + tag) representation; see DESIGN.md "score range"
int check_range(const age_job &j)
{
    const int m = j.match, mm = j.mismatch, go = j.gap_open, ge = j.gap_extend;
    if (m > 63 || m < -63 || mm > 63 || mm < -63) return AGE_ERR_RANGE;
    if (go >= 0 || ge >= 0 || go < -4000 || ge < -4000) return AGE_ERR_RANGE;
    const long k = 2L * (ge - go) - 1;
    const long kabs = k < 0 ? -k : k;
    const long top = 2L * std::max(m, 0) * std::min(j.len1, j.len2);
    if (top + kabs + 2L * (-go) + 2L * (-ge) + 8 > 32767) return AGE_ERR_RANGE;
    return AGE_OK;
}

void build_plan(age_ctx *ctx, const age_job *jobs, size_t n)
{
    ctx->jobs = jobs; ctx->n_jobs = n;
    ctx->job_status.assign(n, AGE_OK);
    ctx->aligners.clear(); ctx->pairs.clear();
    for (size_t i = 0; i < n; ++i) {
        const age_job &j = jobs[i];
        int st = AGE_OK;
        const bool inv = j.flag == F_INV;
        if (!j.seq1 || !j.seq2 || j.len1 <= 0 || j.len2 <= 0) st = AGE_NOT_ALIGNED;        // AGEaligner.cpp:68
        else if (!inv && !single_mode(j.flag)) st = AGE_NOT_ALIGNED;                         // AGEaligner.cpp:69
        else st = check_range(j);
        ctx->job_status[i] = st;
        if (st != AGE_OK) continue;
        if (inv) {
            ctx->aligners.push_back({(int)i, 0, F_INVL, -1});
            ctx->aligners.push_back({(int)i, 1, F_INVR, -1});
        } else ctx->aligners.push_back({(int)i, 0, j.flag, -1});
    }
    // pair aligners of identical shape and scoring.  Selection pairs first: INVL+INVR of one -inv job
    // (AGEaligner.cpp:136-154) and mutual -both partners (age_align.cpp:264-278) share a warp, and only the
    // half that the reference would print is enumerated and traced back.
    struct Key { int32_t v[6]; };
    auto key = [&](int a) {
        const age_job &j = jobs[ctx->aligners[a].job];
        return Key{{j.len1, j.len2, j.match, j.mismatch, j.gap_open, j.gap_extend}};
    };
    auto less = [&](int x, int y) { const Key a = key(x), b = key(y); return std::lexicographical_compare(a.v, a.v + 6, b.v, b.v + 6); };
    auto same = [&](int x, int y) { return memcmp(key(x).v, key(y).v, sizeof(Key)) == 0; };
    auto add_pair = [&](int a0, int a1, int select) {
        HostPair hp{};
        hp.a[0] = a0; hp.a[1] = a1; hp.select = select;
        const age_job &j = jobs[ctx->aligners[a0].job];
        hp.len1 = j.len1; hp.len2 = j.len2;
        hp.nstrips = (j.len2 + STRIP - 1) / STRIP; hp.P = hp.nstrips * STRIP;
        hp.nb = (j.len1 + W - 1) / W; hp.nsteps = hp.nb + 31; hp.nwin = (hp.nsteps + CKS - 1) / CKS;
        hp.k7 = (ctx->aligners[a0].flag != F_INDEL) || (a1 >= 0 && ctx->aligners[a1].flag != F_INDEL);
        const size_t slots = (size_t)hp.nstrips * hp.nsteps;                 // (strip, step) slots of the diagonal-major planes
        hp.in_bytes = 2 * (align_up((size_t)j.len1, 16) + align_up((size_t)j.len2, 16));   // raw sequences of both halves
        hp.scratch_bytes = align_up(slots * W * 2 * 32 * 16, 256) +
                           2 * align_up((size_t)std::max(hp.nstrips - 1, 0) * 3 * (hp.nb + 1) * 16, 256) +
                           2 * align_up(((size_t)hp.P + 2) * 4, 256) + align_up(((size_t)W * hp.nb + 4) * 4, 256) +
                           align_up((size_t)hp.nstrips * hp.nwin * 32 * 4, 256) +
                           2 * align_up((size_t)hp.nstrips * hp.nwin * CKW * 32 * 4, 256);
        hp.scratch_bytes += 2 * align_up(((size_t)hp.nb + 2) * 16, 256) + align_up((size_t)hp.P * 8, 256) + align_up((size_t)hp.P, 256);
        hp.nsets = (select || a1 < 0) ? 1 : 2;
        hp.set_bytes = p2_set_bytes(slots, (size_t)hp.nstrips * hp.nwin, (size_t)j.len2, hp.k7);
        hp.scratch_bytes += hp.nsets * hp.set_bytes;
        hp.cells = (double)j.len1 * j.len2;
        ctx->pairs.push_back(hp);
    };
    std::vector<int> first_of(n, -1);
    for (size_t a = 0; a < ctx->aligners.size(); ++a) if (ctx->aligners[a].role == 0) first_of[ctx->aligners[a].job] = (int)a;
    std::vector<char> used(ctx->aligners.size(), 0);
    for (size_t a = 0; a < ctx->aligners.size(); ++a) {
        const Aligner &al = ctx->aligners[a];
        if (used[a] || al.role != 0) continue;
        const age_job &j = jobs[al.job];
        if (j.flag == F_INV) { add_pair((int)a, (int)a + 1, 1); used[a] = used[a + 1] = 1; continue; }
        const int pj = j.partner;
        if (pj > al.job && (size_t)pj < n && jobs[pj].partner == al.job && ctx->job_status[pj] == AGE_OK &&
            jobs[pj].flag != F_INV && same((int)a, first_of[pj])) {
            add_pair((int)a, first_of[pj], 1); used[a] = used[first_of[pj]] = 1;
        }
    }
    std::vector<int> order;
    for (size_t i = 0; i < ctx->aligners.size(); ++i) if (!used[i]) order.push_back((int)i);
    std::stable_sort(order.begin(), order.end(), less);
    for (size_t i = 0; i < order.size();) {
        if (i + 1 < order.size() && same(order[i + 1], order[i])) { add_pair(order[i], order[i + 1], 0); i += 2; }
        else { add_pair(order[i], -1, 0); i += 1; }
    }
    ctx->outs.assign(ctx->aligners.size(), AlignerOut{});
    ctx->blocks.assign(ctx->aligners.size(), {});
}

uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// Copy the raw sequences of pair `pi` into the pinned host blob at `h` (device address `d`) and fill its task record.
// The column descriptors and row tables are derived from the raw strings on the device (age_pack_kernel).
void pack_pair(age_ctx *ctx, int pi, char *h, char *d, PairTask &T)
{
    const HostPair &hp = ctx->pairs[pi];
    memset(&T, 0, sizeof(T));
    T.len1 = hp.len1; T.len2 = hp.len2; T.nstrips = hp.nstrips; T.P = hp.P;
    T.select = hp.select; T.nb = hp.nb; T.nsteps = hp.nsteps; T.nwin = hp.nwin;
    int sc[2][4] = {{1, -1, -1, -1}, {1, -1, -1, -1}};
    size_t off = 0;
    const char *first_seq1 = nullptr;
    for (int k = 0; k < 2; ++k) {
        T.flag[k] = 0; T.out[k] = -1;
        if (hp.a[k] < 0) continue;
        const Aligner &al = ctx->aligners[hp.a[k]];
        const age_job &j = ctx->jobs[al.job];
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
}

void assign_scratch(const HostPair &hp, char *s, PairTask &T)
{
    const size_t slots = (size_t)hp.nstrips * hp.nsteps;
    const size_t bsz = align_up((size_t)std::max(hp.nstrips - 1, 0) * 3 * (hp.nb + 1) * 16, 256);
    const size_t csz = align_up((size_t)hp.nstrips * hp.nwin * CKW * 32 * 4, 256);
    T.D = (uint32_t *)s; s += align_up(slots * W * 2 * 32 * 16, 256);
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
    T.p2set = s; T.p2stride = (uint32_t)hp.set_bytes; T.k7 = hp.k7;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { dev.err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// stage one sub-batch (pair ids) on a device: pack, copy inputs, lay out scratch
bool stage_on_device(age_ctx *ctx, Device &dev, const std::vector<int> &pair_ids_in)
{
    CK(cudaSetDevice(dev.id));
    // longest pairs first: the persistent sweep kernel hands out pairs in this order, so its tail is made of short ones
    dev.pair_ids = pair_ids_in;
    std::stable_sort(dev.pair_ids.begin(), dev.pair_ids.end(), [&](int a, int b) { return ctx->pairs[a].cells * ctx->pairs[a].nstrips > ctx->pairs[b].cells * ctx->pairs[b].nstrips; });
    const std::vector<int> &pair_ids = dev.pair_ids;
    const size_t np = pair_ids.size();
    dev.tasks.assign(np, PairTask{});
    std::vector<size_t> in_off(np + 1, 0), sc_off(np + 1, 0);
    int n_outs = 0;
    for (size_t i = 0; i < np; ++i) {
        HostPair &hp = ctx->pairs[pair_ids[i]];
        in_off[i + 1] = in_off[i] + hp.in_bytes;
        sc_off[i + 1] = sc_off[i] + hp.scratch_bytes;
        for (int k = 0; k < 2; ++k) if (hp.a[k] >= 0) ctx->aligners[hp.a[k]].out = n_outs++;
    }
    dev.n_outs = n_outs;
    size_t nwt_total = 0;
    for (size_t i = 0; i < np; ++i) { const HostPair &hp = ctx->pairs[pair_ids[i]]; nwt_total += (size_t)hp.nstrips * hp.nwin * hp.nsets; }
    dev.queue_cap = (int)std::min<size_t>(2 * nwt_total + 64, (size_t)1 << 24);
    if (const char *e = getenv("AGE_QUEUE_CAP")) dev.queue_cap = std::max(1, atoi(e));   // test knob: a short work list forces on-demand re-sweeps
    const size_t task_bytes = align_up(np * sizeof(PairTask), 256);
    dev.in_bytes = task_bytes + in_off[np];
    if (!dev.hin.ensure(dev.in_bytes)) { dev.err = "pinned host allocation failed"; return false; }
    if (!dev.in.ensure(dev.in_bytes)) { dev.err = "device allocation failed (inputs)"; return false; }
    if (!dev.scratch.ensure(sc_off[np] + 256)) { dev.err = "device allocation failed (scratch)"; return false; }
    dev.pool_cap = n_outs * 96 + 4096;
    const size_t out_bytes = align_up((size_t)n_outs * sizeof(AlignerOut), 256) + align_up((size_t)dev.pool_cap * sizeof(Block), 256) + 256;
    if (!dev.outbuf.ensure(out_bytes)) { dev.err = "device allocation failed (outputs)"; return false; }
    if (!dev.hout.ensure(out_bytes)) { dev.err = "pinned host allocation failed"; return false; }
    if (!dev.aux.ensure(align_up(np * sizeof(PairHdr), 256) + (size_t)dev.queue_cap * sizeof(uint2))) { dev.err = "device allocation failed (work list)"; return false; }
    parallel_for(ctx->host_threads, np, [&](size_t i) {
        pack_pair(ctx, pair_ids[i], dev.hin.ptr + task_bytes + in_off[i], dev.in.ptr + task_bytes + in_off[i], dev.tasks[i]);
        assign_scratch(ctx->pairs[pair_ids[i]], dev.scratch.ptr + sc_off[i], dev.tasks[i]);
    });
    memcpy(dev.hin.ptr, dev.tasks.data(), np * sizeof(PairTask));
    CK(cudaEventRecord(dev.ev[0], dev.st));
    CK(cudaMemcpyAsync(dev.in.ptr, dev.hin.ptr, dev.in_bytes, cudaMemcpyHostToDevice, dev.st));
    CK(cudaEventRecord(dev.ev[1], dev.st));
    launch_pack((const PairTask *)dev.in.ptr, (int)np, dev.st);
    return true;
}

struct OutLayout { AlignerOut *outs; Block *pool; int32_t *counters; };
OutLayout out_layout(Device &dev, char *base)
{
    OutLayout L;
    L.outs = (AlignerOut *)base;
    L.pool = (Block *)(base + align_up((size_t)dev.n_outs * sizeof(AlignerOut), 256));
    L.counters = (int32_t *)((char *)L.pool + align_up((size_t)dev.pool_cap * sizeof(Block), 256));
    return L;
}

Pass2Params pass2_params(Device &dev, const OutLayout &L)
{
    Pass2Params p{};
    p.pairs = (const PairTask *)dev.in.ptr; p.n_pairs = (int)dev.tasks.size(); p.outs = L.outs;
    p.hdr = (PairHdr *)dev.aux.ptr;
    p.queue = (uint2 *)(dev.aux.ptr + align_up(dev.tasks.size() * sizeof(PairHdr), 256)); p.queue_cap = dev.queue_cap; p.queue_n = L.counters + 12;
    p.pool = L.pool; p.pool_cap = dev.pool_cap; p.pool_used = L.counters + 2; p.next = L.counters + 4;   // next: 8 ints
    return p;
}

bool run_on_device(age_ctx *, Device &dev)
{
    CK(cudaSetDevice(dev.id));
    const int np = (int)dev.tasks.size();
    OutLayout L = out_layout(dev, dev.outbuf.ptr);
    CK(cudaMemsetAsync(L.counters, 0, 64, dev.st));
    CK(cudaEventRecord(dev.ev[2], dev.st));
    int max_strips = 1;
    for (const PairTask &t : dev.tasks) max_strips = std::max(max_strips, (int)t.nstrips);
    launch_sweeps((const PairTask *)dev.in.ptr, np, max_strips, L.counters + 0, dev.st);
    CK(cudaEventRecord(dev.ev[3], dev.st));
    launch_pass2(pass2_params(dev, L), dev.st);
    CK(cudaEventRecord(dev.ev[4], dev.st));
    CK(cudaGetLastError());
    return true;
}

bool fetch_from_device(age_ctx *ctx, Device &dev)
{
    CK(cudaSetDevice(dev.id));
    for (int attempt = 0; attempt < 4; ++attempt) {
        const size_t out_bytes = align_up((size_t)dev.n_outs * sizeof(AlignerOut), 256) + align_up((size_t)dev.pool_cap * sizeof(Block), 256) + 256;
        CK(cudaMemcpyAsync(dev.hout.ptr, dev.outbuf.ptr, out_bytes, cudaMemcpyDeviceToHost, dev.st));
        CK(cudaEventRecord(dev.ev[5], dev.st));
        CK(cudaStreamSynchronize(dev.st));
        OutLayout H = out_layout(dev, dev.hout.ptr);
        if (H.counters[2] <= dev.pool_cap) {
            for (size_t i = 0; i < dev.pair_ids.size(); ++i) {
                const HostPair &hp = ctx->pairs[dev.pair_ids[i]];
                for (int k = 0; k < 2; ++k) {
                    if (hp.a[k] < 0) continue;
                    const Aligner &al = ctx->aligners[hp.a[k]];
                    const AlignerOut &o = H.outs[al.out];
                    memcpy(&ctx->outs[hp.a[k]], &o, offsetof(AlignerOut, bp) + sizeof(o.bp[0]) * (size_t)std::max(o.n_bp, 0));
                    std::vector<age_block> &b = ctx->blocks[hp.a[k]];
                    b.clear();
                    for (int q = 0; q < 4; ++q)
                        for (int x = 0; x < o.n_blk[q]; ++x) {
                            const Block &s = H.pool[o.blk_off[q] + x];
                            b.push_back(age_block{s.start1, s.start2, s.length});
                        }
                }
            }
            ctx->stats.d2h_bytes += out_bytes;
            return true;
        }
        // block pool overflow: enlarge and repeat pass 2 (the sweep outputs are still resident)
        dev.pool_cap = H.counters[2] + 4096;
        const size_t nb = align_up((size_t)dev.n_outs * sizeof(AlignerOut), 256) + align_up((size_t)dev.pool_cap * sizeof(Block), 256) + 256;
        if (!dev.outbuf.ensure(nb) || !dev.hout.ensure(nb)) { dev.err = "allocation failed (block pool)"; return false; }
        OutLayout L = out_layout(dev, dev.outbuf.ptr);
        CK(cudaMemsetAsync(L.counters, 0, 64, dev.st));
        launch_pass2(pass2_params(dev, L), dev.st);
    }
    dev.err = "block pool overflow";
    return false;
}

void add_timing(age_ctx *ctx, Device &dev)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, dev.ev[0], dev.ev[1]) == cudaSuccess) ctx->stats.h2d_ms += ms;
    if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[3]) == cudaSuccess) ctx->stats.fill_ms += ms;
    if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[4]) == cudaSuccess) ctx->stats.kernel_ms += ms;
    if (cudaEventElapsedTime(&ms, dev.ev[4], dev.ev[5]) == cudaSuccess) ctx->stats.d2h_ms += ms;
}

// split the pairs of the plan into per-device sub-batches that fit the scratch budget
std::vector<std::vector<std::vector<int>>> make_subbatches(age_ctx *ctx)
{
    const int nd = (int)ctx->devs.size();
    std::vector<std::vector<std::vector<int>>> res(nd);
    std::vector<size_t> budget(nd, 0);
    auto query_budget = [&](int d) {             // cudaMemGetInfo costs milliseconds: only asked when a batch outgrows the arenas
        cudaSetDevice(ctx->devs[d].id);
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        fr += ctx->devs[d].scratch.cap + ctx->devs[d].in.cap + ctx->devs[d].outbuf.cap + ctx->devs[d].aux.cap;
        budget[d] = (size_t)(0.85 * (double)fr);
    };
    // longest-processing-time-first deal across devices
    std::vector<int> order(ctx->pairs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::vector<std::vector<int>> per_dev(nd);
    if (nd == 1) per_dev[0] = order;
    else {
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ctx->pairs[a].cells > ctx->pairs[b].cells; });
        std::vector<double> load(nd, 0.0);
        for (int pi : order) {
            int best = 0;
            for (int d = 1; d < nd; ++d) if (load[d] < load[best]) best = d;
            per_dev[best].push_back(pi); load[best] += ctx->pairs[pi].cells;
        }
        for (auto &v : per_dev) std::sort(v.begin(), v.end());
    }
    for (int d = 0; d < nd; ++d) {
        size_t sc = 256, in = align_up(per_dev[d].size() * sizeof(PairTask), 256);
        for (int pi : per_dev[d]) { sc += ctx->pairs[pi].scratch_bytes; in += ctx->pairs[pi].in_bytes; }
        if (sc <= ctx->devs[d].scratch.cap && in <= ctx->devs[d].in.cap) {        // fits what is already allocated
            if (!per_dev[d].empty()) res[d].push_back(per_dev[d]);
            continue;
        }
        query_budget(d);
        std::vector<int> cur; size_t used = 0;
        for (int pi : per_dev[d]) {
            const size_t need = ctx->pairs[pi].scratch_bytes + ctx->pairs[pi].in_bytes + 4096;
            if (!cur.empty() && used + need > budget[d]) { res[d].push_back(cur); cur.clear(); used = 0; }
            cur.push_back(pi); used += need;
        }
        if (!cur.empty()) res[d].push_back(cur);
    }
    return res;
}

void account(age_ctx *ctx)
{
    ctx->stats.n_aligners += ctx->aligners.size();
    for (const HostPair &hp : ctx->pairs)
        for (int k = 0; k < 2; ++k) if (hp.a[k] >= 0) ctx->stats.cell_updates += (uint64_t)(2.0 * hp.cells);
}

void assemble(age_ctx *ctx, age_result *out, size_t n)
{
    std::vector<int> main_of(n, -1), aux_of(n, -1);
    for (size_t a = 0; a < ctx->aligners.size(); ++a) {
        const Aligner &al = ctx->aligners[a];
        (al.role ? aux_of : main_of)[al.job] = (int)a;
    }
    for (size_t i = 0; i < n; ++i) {
        age_result &r = out[i];
        memset(&r, 0, offsetof(age_result, bp));          // header only: bp rows beyond n_bp are not part of the result
        r.alt_index = -1; r.n_left = r.n_right = r.n_alt_left = r.n_alt_right = 0;
        r.left = r.right = r.alt_left = r.alt_right = nullptr;
        r.status = ctx->job_status[i];
        r.aux_score = -1;
        if (r.status != AGE_OK) continue;
        const int m = main_of[i], x = aux_of[i];
        const AlignerOut &om = ctx->outs[m];
        r.main_score = om.score;
        int pick = m;
        if (x >= 0) {
            r.aux_score = ctx->outs[x].score;
            if (ctx->outs[x].score > om.score) { pick = x; r.used_aux = 1; }       // AGEaligner.cpp:151,181-184
        }
        const AlignerOut &o = ctx->outs[pick];
        if (o.status != 0) { r.status = AGE_ERR_CUDA; continue; }
        r.score = o.score; r.flag = ctx->aligners[pick].flag;
        if (o.n_bp < 0) continue;                                  // lost the -both selection: score only
        r.detail = 1;
        r.n_bp = o.n_bp;
        memcpy(r.bp, o.bp, sizeof(r.bp[0]) * (size_t)std::max(o.n_bp, 0));
        r.alt_index = o.alt_index;
        r.n_left = o.n_blk[0]; r.n_right = o.n_blk[1]; r.n_alt_left = o.n_blk[2]; r.n_alt_right = o.n_blk[3];
        const age_block *b = ctx->blocks[pick].data();
        r.left = b; r.right = b + r.n_left; r.alt_left = r.right + r.n_right; r.alt_right = r.alt_left + r.n_alt_left;
    }
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
    for (Device &d : ctx->devs) {
        bool ok = cudaSetDevice(d.id) == cudaSuccess && cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking) == cudaSuccess;
        for (int e = 0; ok && e < 6; ++e) ok = cudaEventCreate(&d.ev[e]) == cudaSuccess;
        if (!ok) { g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError()); age_destroy(ctx); return nullptr; }
    }
    if (g_hprof) fprintf(stderr, "[host prof] age_create %.1f ms\n", wall_ms() - t_create);
    return ctx;
}

extern "C" void age_destroy(age_ctx *ctx)
{
    if (!ctx) return;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        if (d.st) cudaStreamSynchronize(d.st);
        d.in.release(); d.scratch.release(); d.outbuf.release(); d.aux.release(); d.hin.release(); d.hout.release();
        for (auto &e : d.ev) if (e) cudaEventDestroy(e);
        if (d.st) cudaStreamDestroy(d.st);
    }
    delete ctx;
}

extern "C" const char *age_last_error(const age_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static int run_plan(age_ctx *ctx, age_result *out, size_t n)
{
    static const bool hprof = getenv("AGE_HOST_PROF") != nullptr;
    double tp[6] = {0, 0, 0, 0, 0, 0};
    const double tsub = wall_ms();
    auto sub = make_subbatches(ctx);
    tp[3] = wall_ms() - tsub;
    const int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, 0);
    std::vector<age_stats> dstat(nd);
    auto worker = [&](int d) {
        Device &dev = ctx->devs[d];
        for (auto &ids : sub[d]) {
            if (!stage_on_device(ctx, dev, ids) || !run_on_device(ctx, dev) || !fetch_from_device(ctx, dev)) { rc[d] = AGE_ERR_CUDA; return; }
            float ms = 0;
            dstat[d].h2d_bytes += dev.in_bytes;
            if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[3]) == cudaSuccess) dstat[d].fill_ms += ms;
            if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[4]) == cudaSuccess) dstat[d].kernel_ms += ms;
        }
    };
    // timing is accumulated after the join (stats are not thread-safe)
    if (nd == 1) {
        Device &dev = ctx->devs[0];
        for (auto &ids : sub[0]) {
            double t0 = wall_ms();
            bool ok = stage_on_device(ctx, dev, ids);
            if (hprof) { cudaStreamSynchronize(dev.st); tp[0] += wall_ms() - t0; t0 = wall_ms(); }
            ok = ok && run_on_device(ctx, dev);
            if (hprof) { cudaStreamSynchronize(dev.st); tp[1] += wall_ms() - t0; t0 = wall_ms(); }
            ok = ok && fetch_from_device(ctx, dev);
            if (hprof) tp[2] += wall_ms() - t0;
            if (!ok) { rc[0] = AGE_ERR_CUDA; break; }
            ctx->stats.h2d_bytes += dev.in_bytes; ctx->stats.launches += 8; ctx->stats.fill_launches += 1;
            add_timing(ctx, dev);
        }
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; ++d) th.emplace_back(worker, d);
        for (auto &t : th) t.join();
        for (int d = 0; d < nd; ++d) {      // devices run concurrently: report the slowest one
            ctx->stats.launches += 8 * sub[d].size(); ctx->stats.fill_launches += sub[d].size();
            ctx->stats.h2d_bytes += dstat[d].h2d_bytes;
            ctx->stats.kernel_ms = std::max(ctx->stats.kernel_ms, dstat[d].kernel_ms);
            ctx->stats.fill_ms = std::max(ctx->stats.fill_ms, dstat[d].fill_ms);
        }
    }
    for (int d = 0; d < nd; ++d)
        if (rc[d]) { ctx->err = ctx->devs[d].err; return rc[d]; }
    account(ctx);
    double t0 = wall_ms();
    assemble(ctx, out, n);
    if (hprof) fprintf(stderr, "[host prof] sub-batching %.2f ms, stage(pack+H2D) %.2f ms, kernels %.2f ms, fetch(D2H+copy) %.2f ms, assemble %.2f ms\n", tp[3], tp[0], tp[1], tp[2], wall_ms() - t0);
    return 0;
}

extern "C" int age_align_batch(age_ctx *ctx, const age_job *jobs, size_t n, age_result *out)
{
    if (!ctx || (!jobs && n) || (!out && n)) return AGE_ERR_ARG;
    ctx->stats = age_stats{};
    ctx->staged = false;
    const double tb = wall_ms();
    build_plan(ctx, jobs, n);
    const double tplan = wall_ms() - tb;
    const int rc = run_plan(ctx, out, n);
    if (getenv("AGE_HOST_PROF")) fprintf(stderr, "[host prof] plan %.2f ms, age_align_batch total %.2f ms\n", tplan, wall_ms() - tb);
    return rc;
}

extern "C" int age_stage(age_ctx *ctx, const age_job *jobs, size_t n)
{
    if (!ctx || (!jobs && n)) return AGE_ERR_ARG;
    ctx->stats = age_stats{};
    ctx->staged = false;
    build_plan(ctx, jobs, n);
    auto sub = make_subbatches(ctx);
    for (size_t d = 0; d < ctx->devs.size(); ++d) {
        if (sub[d].size() > 1) { ctx->err = "batch does not fit in device memory in one piece; use age_align_batch"; return AGE_ERR_ARG; }
        Device &dev = ctx->devs[d];
        if (sub[d].empty()) { dev.pair_ids.clear(); dev.tasks.clear(); dev.n_outs = 0; continue; }
        if (!stage_on_device(ctx, dev, sub[d][0])) { ctx->err = dev.err; return AGE_ERR_CUDA; }
        ctx->stats.h2d_bytes += dev.in_bytes;
    }
    for (Device &dev : ctx->devs) { cudaSetDevice(dev.id); cudaStreamSynchronize(dev.st); }
    ctx->staged = true;
    return 0;
}

extern "C" int age_run_staged(age_ctx *ctx)
{
    if (!ctx || !ctx->staged) return AGE_ERR_ARG;
    for (Device &dev : ctx->devs) {
        if (dev.tasks.empty()) continue;
        if (!run_on_device(ctx, dev)) { ctx->err = dev.err; return AGE_ERR_CUDA; }
    }
    ctx->stats.kernel_ms = 0; ctx->stats.fill_ms = 0;
    for (Device &dev : ctx->devs) {
        if (dev.tasks.empty()) continue;
        cudaSetDevice(dev.id);
        if (cudaStreamSynchronize(dev.st) != cudaSuccess) { ctx->err = cudaGetErrorString(cudaGetLastError()); return AGE_ERR_CUDA; }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[4]) == cudaSuccess) ctx->stats.kernel_ms = std::max(ctx->stats.kernel_ms, (double)ms);
        if (cudaEventElapsedTime(&ms, dev.ev[2], dev.ev[3]) == cudaSuccess) ctx->stats.fill_ms = std::max(ctx->stats.fill_ms, (double)ms);
    }
    ctx->stats.launches += 8; ctx->stats.fill_launches += 1;
    return 0;
}

extern "C" int age_fetch(age_ctx *ctx, age_result *out, size_t n)
{
    if (!ctx || !ctx->staged || n != ctx->n_jobs || (!out && n)) return AGE_ERR_ARG;
    for (Device &dev : ctx->devs) {
        if (dev.tasks.empty()) continue;
        if (!fetch_from_device(ctx, dev)) { ctx->err = dev.err; return AGE_ERR_CUDA; }
    }
    ctx->stats.n_aligners = 0; ctx->stats.cell_updates = 0;
    account(ctx);
    assemble(ctx, out, n);
    return 0;
}

extern "C" int age_get_stats(const age_ctx *ctx, age_stats *st)
{
    if (!ctx || !st) return AGE_ERR_ARG;
    *st = ctx->stats;
    return 0;
}

extern "C" const char *age_version(void) { return "age-b200 0.1 (sm_100a)"; }
