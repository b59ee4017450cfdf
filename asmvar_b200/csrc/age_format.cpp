// age_format.cpp -- host-side `.age` record formatter.
//
// Produces byte-for-byte the text that the reference's AGEaligner::printAlignment()
// (src/AsmvarAGE/src/AGEaligner.cpp:179-380) and AliFragment::printAlignment()
// (AGEaligner.hh:230-272) write to stdout, from the compact result the GPU returns
// (score, breakpoints, diagonal block lists) plus the original strings.
// No iostream: at GPU rates the formatter is on the critical path, so it appends into a
// flat byte buffer.
#include "age_b200.h"
#include "age_host.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

namespace age {

char complement(char c)
{   // Sequence.hh:138-150
    switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 't': return 'a'; case 'g': return 'c';
    case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C';
    case 'u': return 'a'; case 'U': return 'A';
    default:  return c;
    }
}

static inline int up(char c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// Nucleotide classes of Sequence::sameNuc (Sequence.hh:182-188): the five letters it accepts, case-insensitive; 0 = anything
// else, which never matches (not even itself).
struct NucTable {
    unsigned char v[256];
    constexpr NucTable() : v{}
    {
        v[(unsigned char)'A'] = v[(unsigned char)'a'] = 1; v[(unsigned char)'C'] = v[(unsigned char)'c'] = 2;
        v[(unsigned char)'G'] = v[(unsigned char)'g'] = 3; v[(unsigned char)'T'] = v[(unsigned char)'t'] = 4;
        v[(unsigned char)'U'] = v[(unsigned char)'u'] = 5;
    }
};
static constexpr NucTable NUC{};
static inline bool same_nuc_fast(char a, char b) { const unsigned char x = NUC.v[(unsigned char)a]; return x != 0 && x == NUC.v[(unsigned char)b]; }

bool same_nuc(char a, char b) { return same_nuc_fast(a, b); }

static inline bool is_gap(char c) { return c == ' ' || c == '-'; }   // Sequence.hh:175-180

// ---- flat output buffer --------------------------------------------------------------------
struct Out {
    char *buf; size_t cap; size_t n;
    void put(char c) { if (n + 1 < cap) buf[n] = c; n++; }
    void put(const char *s, size_t len)
    {
        if (n + len < cap) memcpy(buf + n, s, len);
        else for (size_t i = 0; i < len; i++) if (n + i + 1 < cap) buf[n + i] = s[i];
        n += len;
    }
    void put(const char *s) { put(s, strlen(s)); }
    void fill(char c, size_t len)
    {
        if (n + len < cap) memset(buf + n, c, len);
        else for (size_t i = 0; i < len; i++) if (n + i + 1 < cap) buf[n + i] = c;
        n += len;
    }
    void num(long v, int width = 0)
    {   // operator<<(int) with setw(width), right aligned, fill ' '
        char tmp[24]; int k = 0; bool neg = v < 0; unsigned long u = neg ? 0ul - (unsigned long)v : (unsigned long)v;
        do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) tmp[k++] = '-';
        for (int i = k; i < width; i++) put(' ');
        while (k) put(tmp[--k]);
    }
};

// Column width of the coordinate pairs in the header lines (the reference's calcWidth, AGEaligner.cpp:382-390): the
// larger digit count of the two values, where 0 has no digit and a negative value counts the digits of -(v + 1).
static int digit_count(long v)
{
    unsigned long u = v < 0 ? (unsigned long)(-(v + 1)) : (unsigned long)v;
    int d = 0;
    for (; u; u /= 10) ++d;
    return d;
}
static int calc_width(int v1, int v2) { return std::max(digit_count(v1), digit_count(v2)); }

// One aligned fragment (class AliFragment, AGEaligner.hh:179-273)
struct Frag {
    std::string a1, a2;
    int start1, start2, end1, end2;
};

struct SeqHdr { const char *seq; int len; int start; bool reverse; const unsigned char *nuc = nullptr; };     // nuc: classes of NUC per position (optional)

static inline char at(const SeqHdr &s, long i) { return (i >= 0 && i < s.len) ? s.seq[i] : '\0'; }

// classes of every base of a sequence for the breakpoint identities (thread-local storage: no allocation per record)
static void classify(SeqHdr &s, std::vector<unsigned char> &store)
{
    store.resize((size_t)std::max(s.len, 1));
    for (int i = 0; i < s.len; i++) store[(size_t)i] = NUC.v[(unsigned char)s.seq[i]];
    s.nuc = store.data();
}

// findAlignment (AGEaligner.cpp:670-750): blocks -> gapped strings + coordinates
static void build_frags(const SeqHdr &s1, const SeqHdr &s1rc, const SeqHdr &s2, int flag,
                        const age_block *left, int n_left, const age_block *right, int n_right,
                        std::vector<Frag> &frags)
{
    frags.clear();
    const age_block *bs[2] = {left, right};
    const int nb[2] = {n_left, n_right};
    for (int b = 0; b < 2; b++) {
        if (nb[b] <= 0 || !bs[b]) continue;
        bool use_rc1 = ((flag & AGE_INVL_FLAG) && b == 1) || ((flag & AGE_INVR_FLAG) && b == 0);
        const SeqHdr &q1 = use_rc1 ? s1rc : s1;
        int inc1 = q1.reverse ? -1 : 1, inc2 = s2.reverse ? -1 : 1;
        Frag f;
        const age_block *bl = bs[b];
        f.start1 = q1.start + inc1 * (bl[0].start1 - 1);
        f.start2 = s2.start + inc2 * (bl[0].start2 - 1);
        f.end1 = f.end2 = 0;
        int ind1 = bl[0].start1 - 1, ind2 = bl[0].start2 - 1;
        // columns of the fragment: every base of either sequence between the first and the last block is one
        size_t cols = 0;
        {
            int i1 = ind1, i2 = ind2;
            for (int k = 0; k < nb[b]; k++) {
                cols += (size_t)std::max(0, bl[k].start1 - 1 - i1) + (size_t)std::max(0, bl[k].start2 - 1 - i2) + (size_t)std::max(0, bl[k].length);
                i1 = std::max(i1, bl[k].start1 - 1) + std::max(0, bl[k].length); i2 = std::max(i2, bl[k].start2 - 1) + std::max(0, bl[k].length);
            }
        }
        f.a1.resize(cols); f.a2.resize(cols);
        char *p1 = &f.a1[0], *p2 = &f.a2[0];
        size_t w = 0;
        // (blocks inside both sequences: plain copies; anything else goes through at(), which yields '\0' outside)
        auto copy1 = [&](char *dst, const SeqHdr &q, int from, int len) {
            if (from >= 0 && from + len <= q.len) memcpy(dst, q.seq + from, (size_t)len);
            else for (int i = 0; i < len; i++) dst[i] = at(q, (long)from + i);
        };
        for (int k = 0; k < nb[b]; k++) {
            f.end1 = q1.start + inc1 * (bl[k].start1 + bl[k].length - 2);
            f.end2 = s2.start + inc2 * (bl[k].start2 + bl[k].length - 2);
            int st1 = bl[k].start1 - 1, st2 = bl[k].start2 - 1;
            if (ind1 < st1) { const int g = st1 - ind1; copy1(p1 + w, q1, ind1, g); memset(p2 + w, '-', (size_t)g); w += (size_t)g; ind1 = st1; }
            if (ind2 < st2) { const int g = st2 - ind2; memset(p1 + w, '-', (size_t)g); copy1(p2 + w, s2, ind2, g); w += (size_t)g; ind2 = st2; }
            const int len = std::max(0, bl[k].length);
            copy1(p1 + w, q1, ind1, len); copy1(p2 + w, s2, ind2, len);
            w += (size_t)len; ind1 += len; ind2 += len;
        }
        f.a1.resize(w); f.a2.resize(w);
        if (!f.a1.empty() && f.a1.size() == f.a2.size()) frags.push_back(std::move(f));
    }
}

// The excised region between two consecutive fragments in first-sequence coordinates (semantics of getExcisedRange,
// AGEaligner.cpp:156-177).  Per mode, which fragment end opens the region and whether that end itself belongs to it:
//   INDEL  (f.end1, nx.start1) both exclusive      TDUPLICATION  [nx.start1, f.end1]: the duplicated stretch, ends swapped
//   INVL   (f.end1, nx.start1] left exclusive      INVR          [f.end1, nx.start1) right exclusive
static bool excised_range(int flag, bool rev1, const Frag &f, const Frag &nx, int &s, int &e, int add_s = 0, int add_e = 0)
{
    const int step = rev1 ? -1 : 1;
    const int mode = (flag & AGE_INDEL_FLAG) ? 0 : (flag & AGE_TDUPLICATION_FLAG) ? 1 : (flag & AGE_INVL_FLAG) ? 2 : (flag & AGE_INVR_FLAG) ? 3 : -1;
    if (mode < 0) return false;
    if (mode == 1) { s = nx.start1; e = f.end1; }
    else {
        s = f.end1 + ((mode == 0 || mode == 2) ? step : 0);
        e = nx.start1 - ((mode == 0 || mode == 3) ? step : 0);
    }
    s += step * add_s;
    e += step * add_e;
    return true;
}

// Longest m <= n such that the m bases starting at `head` equal the m bases ending at `tail` (Sequence::sameNuc: both
// valid nucleotides, case-insensitive), by the prefix function of  head[0..n) # tail-n+1..tail  -- one pass instead of the
// reference's restart per candidate length.  Bases that are not nucleotides never match, not even themselves: they get
// different codes on the two sides of the separator, and inside a matched prefix there are none, so the prefix function
// is only ever consulted on proper nucleotides.
static int head_tail_overlap(const SeqHdr &q, long head, long tail, int n)
{
    if (n <= 0) return 0;
    // Nearly every candidate length fails at its first or second base, so the lengths are simply tried from the longest down
    // (what the reference does), on the nucleotide classes of the sequence (SeqHdr::nuc: one table look-up per base, made once
    // per record; then plain byte compares) -- but only with a budget of comparisons: a repetitive stretch that makes many lengths fail
    // late falls through to the prefix function below, which is linear whatever the sequence.
    if (q.nuc && head >= 0 && tail - n + 1 >= 0 && head + n <= q.len && tail < q.len) {
        const unsigned char *hd = q.nuc + head, *tl = q.nuc + tail;
        const unsigned char h0 = hd[0];
        if (h0 == 0) return 0;                                                                 // no length can match at its first base
        long budget = 8L * n + 64;
        int m = n;
        for (; m >= 1 && budget > 0; --m) {
            const unsigned char *t0 = tl - m + 1;
            if (t0[0] != h0) continue;
            int k = 1;
            while (k < m && hd[k] != 0 && hd[k] == t0[k]) ++k;
            budget -= k;
            if (k == m) return m;
        }
        if (m < 1) return 0;
    }
    auto norm = [&](long i, char bad) {
        const unsigned char c = NUC.v[(unsigned char)at(q, i)];
        return c ? (char)(c + 8) : bad;
    };
    thread_local std::vector<char> t;
    thread_local std::vector<int> pf;
    t.assign((size_t)2 * n + 1, 0);
    for (int i = 0; i < n; i++) { t[i] = norm(head + i, 1); t[n + 1 + i] = norm(tail - n + 1 + i, 2); }
    t[n] = 3;
    pf.assign(t.size(), 0);
    for (size_t i = 1; i < t.size(); i++) {
        int k = pf[i - 1];
        while (k > 0 && t[i] != t[k]) k = pf[k - 1];
        pf[i] = k + (t[i] == t[k] ? 1 : 0);
    }
    return pf.back();
}

// "Identity outside breakpoints" (calcOutsideIdentity, AGEaligner.cpp:392-409): the longest stretch that ends at bp1 and is
// repeated starting at bp2 (0-based positions), at most as long as the room on either side.
static int outside_identity(const SeqHdr &q, int bp1, int bp2)
{
    if (bp1 >= bp2) return 0;
    return head_tail_overlap(q, bp2, bp1, std::min(bp1 + 1, q.len - bp2));
}

// "Identity inside breakpoints" (calcInsideIdentity, AGEaligner.cpp:411-427): the longest stretch that starts at bp1 and is
// repeated ending at bp2, at most half of the region.
static int inside_identity(const SeqHdr &q, int bp1, int bp2)
{
    if (bp1 >= bp2) return 0;
    return head_tail_overlap(q, bp1, bp2, (bp2 - bp1 + 1) / 2);
}

// "Identity at breakpoints" (calcIdentity, AGEaligner.cpp:429-447): how far the sequence keeps agreeing with itself shifted
// by bp2 - bp1, to the left of bp1 (left <= 0) and from bp1 on to the right (right >= 0).
static int at_identity(const SeqHdr &q, int bp1, int bp2, int &left, int &right)
{
    if (bp1 >= bp2) return 0;
    const long shift = (long)bp2 - bp1;
    int down = 0, up = 0;
    for (long i = bp1 - 1; i >= 0 && same_nuc_fast(at(q, i), at(q, i + shift)); --i) ++down;
    for (long i = bp1; i < q.len && same_nuc_fast(at(q, i), at(q, i + shift)); ++i) ++up;
    left = -down; right = up;
    return up + down;
}

static void range_line(Out &o, const char *label, int len, int s, int e)
{
    o.put(label); o.num(len, 9); o.put(" nucs");
    if (len > 0) { o.put(" ["); o.num(s); o.put(','); o.num(e); o.put(']'); }
    o.put('\n');
}

// the two lines printed per excised region (AGEaligner.cpp:270-284 and :290-304)
static void excised_lines(Out &o, int flag, bool rev1, bool rev2, const std::vector<Frag> &fr)
{
    int inc1 = rev1 ? -1 : 1, inc2 = rev2 ? -1 : 1;
    for (size_t i = 0; i + 1 < fr.size(); i++) {
        int s, e;
        if (!excised_range(flag, rev1, fr[i], fr[i + 1], s, e)) continue;
        range_line(o, " first  seq => ", std::abs(s - e - inc1), s, e);
        s = fr[i].end2 + inc2;
        e = fr[i + 1].start2 - inc2;
        range_line(o, " second seq => ", std::abs(s - e - inc2), s, e);
    }
}

// AliFragment::printAlignment (AGEaligner.hh:230-272)
static void print_frag(Out &o, const Frag &f)
{
    const int WIDTH = 60, MARGIN = 9;
    int inc1 = f.end1 < f.start1 ? -1 : 1;
    int inc2 = f.end2 < f.start2 ? -1 : 1;
    int n = (int)(f.a1.size() < f.a2.size() ? f.a1.size() : f.a2.size());
    int ind1 = f.start1, ind2 = f.start2;
    char match[WIDTH];
    for (int i = 0; i < n; i += WIDTH) {
        int st1 = ind1, st2 = ind2;
        int w = n - i < WIDTH ? n - i : WIDTH;
        const char *a1 = f.a1.data() + i, *a2 = f.a2.data() + i;
        int nuc1 = 0, nuc2 = 0;
        for (int j = 0; j < w; j++) {
            if (same_nuc_fast(a1[j], a2[j])) match[j] = '|';
            else if (!is_gap(a1[j]) && !is_gap(a2[j])) match[j] = '.';
            else match[j] = ' ';
            if (!is_gap(a1[j])) { nuc1++; ind1 += inc1; }
            if (!is_gap(a2[j])) { nuc2++; ind2 += inc2; }
        }
        o.put('\n');
        if (nuc1 > 0) { o.num(st1, MARGIN); o.put(' '); o.put(a1, w); o.put(' '); o.num(ind1 - inc1); o.put('\n'); }
        else { o.fill(' ', MARGIN + 1); o.put(a1, w); o.put('\n'); }
        o.fill(' ', MARGIN + 1);
        o.put(match, w); o.put('\n');
        if (nuc2 > 0) { o.num(st2, MARGIN); o.put(' '); o.put(a2, w); o.put(' '); o.num(ind2 - inc2); o.put('\n'); }
        else { o.fill(' ', MARGIN + 1); o.put(a2, w); o.put('\n'); }
    }
}

size_t format_alignment(const age_seqview *v1, const age_seqview *v2, const age_job *job,
                        const age_result *res, char *buf, size_t cap)
{
    Out o{buf, cap, 0};
    const int flag = res->flag;
    SeqHdr s1{v1->seq, v1->len, v1->start, v1->reverse != 0};
    SeqHdr s2{v2->seq, v2->len, v2->start, v2->reverse != 0};
    // _s1_rc = clone + revcom (AGEaligner.cpp:20,27; Sequence.hh:190-198); only INVL/INVR read it
    std::string rc;
    SeqHdr s1rc = s1;
    if (flag & (AGE_INVL_FLAG | AGE_INVR_FLAG)) {
        rc.resize((size_t)s1.len);
        for (int i = 0; i < s1.len; i++) rc[(size_t)i] = complement(s1.seq[s1.len - 1 - i]);
        s1rc.seq = rc.data();
        s1rc.start = s1.reverse ? s1.start - (s1.len - 1) : s1.start + (s1.len - 1);
        s1rc.reverse = !s1.reverse;
    }

    o.put("\nMATCH = ");    o.num(job->match);
    o.put(", MISMATCH = "); o.num(job->mismatch);
    o.put(", GAP OPEN = "); o.num(job->gap_open);
    o.put(", GAP EXTEND = "); o.num(job->gap_extend);
    if (flag & AGE_INDEL_FLAG)             o.put(", INDEL");
    else if (flag & AGE_INVL_FLAG)         o.put(", INVERSION");
    else if (flag & AGE_INVR_FLAG)         o.put(", INVERSION");
    else if (flag & AGE_TDUPLICATION_FLAG) o.put(", TDUPLICATION");
    o.put("\n\n");

    int inc1 = s1.reverse ? -1 : 1, inc2 = s2.reverse ? -1 : 1;
    int b1 = s1.start, b2 = s2.start;
    int e1 = b1 + inc1 * (s1.len - 1), e2 = b2 + inc2 * (s2.len - 1);
    int wb = calc_width(b1, b2), we = calc_width(e1, e2);
    o.put("First  seq ["); o.num(b1, wb); o.put(','); o.num(e1, we); o.put("] => ");
    o.num(s1.len, 9); o.put(" nucs '"); o.put(v1->name ? v1->name : ""); o.put("'\n");
    o.put("Second seq ["); o.num(b2, wb); o.put(','); o.num(e2, we); o.put("] => ");
    o.num(s2.len, 9); o.put(" nucs '"); o.put(v2->name ? v2->name : ""); o.put("'\n\n");

    std::vector<Frag> frags;
    build_frags(s1, s1rc, s2, flag, res->left, res->n_left, res->right, res->n_right, frags);
    const int n_frg = (int)frags.size();
    std::vector<int> n_ali(n_frg + 1, 0), n_id(n_frg + 1, 0), n_gap(n_frg + 1, 0);
    for (int k = 0; k < n_frg; k++) {   // AliFragment::countAligned (AGEaligner.hh:218-228)
        const Frag &f = frags[k];
        for (size_t i = 0; i < f.a1.size(); i++) {
            n_ali[k + 1]++;
            if (same_nuc_fast(f.a1[i], f.a2[i])) n_id[k + 1]++;
            if (is_gap(f.a1[i]) || is_gap(f.a2[i])) n_gap[k + 1]++;
        }
        n_ali[0] += n_ali[k + 1]; n_id[0] += n_id[k + 1]; n_gap[0] += n_gap[k + 1];
    }
    int identic = 0, gap = 0;
    if (n_ali[0] > 0) {
        identic = (int)(100. * n_id[0] / n_ali[0] + 0.5);
        gap     = (int)(100. * n_gap[0] / n_ali[0] + 0.5);
    }
    // res->score here is the _max of the printed aligner
    int printed_max = res->used_aux ? res->aux_score : res->main_score;
    o.put("Score:   "); o.num(printed_max, 9); o.put('\n');
    o.put("Aligned: "); o.num(n_ali[0], 9); o.put("        nucs\n");
    o.put("Identic: "); o.num(n_id[0], 9); o.put(" ("); o.num(identic, 3); o.put("%) nucs");
    if (n_frg > 1) {
        o.put(" =>");
        for (int k = 1; k <= n_frg; k++) {
            o.put(' '); o.num(n_id[k], 9); o.put(" (");
            o.num((int)(100. * n_id[k] / n_ali[k] + 0.5), 3); o.put("%)");
        }
    }
    o.put("\nGaps:    "); o.num(n_gap[0], 9); o.put(" ("); o.num(gap, 3); o.put("%) nucs\n\n");

    if (n_ali[0] > 0) {
        o.put("Alignment:\n");
        o.put(" first  seq =>  ");
        for (int k = 0; k < n_frg; k++) {
            if (k) o.put(" EXCISED REGION ");
            int ws = calc_width(frags[k].start1, frags[k].start2), wend = calc_width(frags[k].end1, frags[k].end2);
            o.put('['); o.num(frags[k].start1, ws); o.put(','); o.num(frags[k].end1, wend); o.put(']');
        }
        o.put("\n second seq =>  ");
        for (int k = 0; k < n_frg; k++) {
            if (k) o.put(" EXCISED REGION ");
            int ws = calc_width(frags[k].start1, frags[k].start2), wend = calc_width(frags[k].end1, frags[k].end2);
            o.put('['); o.num(frags[k].start2, ws); o.put(','); o.num(frags[k].end2, wend); o.put(']');
        }
        o.put('\n');
    }

    if (n_frg > 1) {
        thread_local std::vector<unsigned char> cls1, cls2;
        classify(s1, cls1); classify(s2, cls2);
        o.put("\nEXCISED REGION(S):\n");
        excised_lines(o, flag, s1.reverse, s2.reverse, frags);

        if (res->n_bp > 1) { o.put("ALTERNATIVE REGION(S): "); o.num(res->n_bp - 1); o.put('\n'); }
        if (res->alt_index >= 1) {
            std::vector<Frag> alt;
            build_frags(s1, s1rc, s2, flag, res->alt_left, res->n_alt_left, res->alt_right, res->n_alt_right, alt);
            excised_lines(o, flag, s1.reverse, s2.reverse, alt);
        }

        o.put("\nIdentity at breakpoints: \n");
        for (int k = 0; k + 1 < n_frg; k++) {
            int s, e, left = 0, right = 0;
            if (!excised_range(flag, s1.reverse, frags[k], frags[k + 1], s, e, 0, 1)) continue;
            int bp1 = inc1 * (s - s1.start), bp2 = inc1 * (e - s1.start);
            int n_hom = at_identity(s1, bp1, bp2, left, right);
            o.put(" first  seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s + inc1 * left); o.put(','); o.num(s + inc1 * (right - 1)); o.put("] to");
                o.put(" ["); o.num(e + inc1 * left); o.put(','); o.num(e + inc1 * (right - 1)); o.put(']');
            }
            o.put('\n');
            s = frags[k].end2 + inc2; e = frags[k + 1].start2;
            bp1 = inc2 * (s - s2.start); bp2 = inc2 * (e - s2.start);
            left = right = 0;
            n_hom = at_identity(s2, bp1, bp2, left, right);
            o.put(" second seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s + inc2 * left); o.put(','); o.num(s + inc2 * (right - 1)); o.put("] to");
                o.put(" ["); o.num(e + inc2 * left); o.put(','); o.num(e + inc2 * (right - 1)); o.put(']');
            }
            o.put('\n');
        }

        o.put("Identity outside breakpoints: \n");
        for (int k = 0; k + 1 < n_frg; k++) {
            int s, e;
            if (!excised_range(flag, s1.reverse, frags[k], frags[k + 1], s, e, -1, 1)) continue;
            int bp1 = inc1 * (s - s1.start), bp2 = inc1 * (e - s1.start);
            int n_hom = outside_identity(s1, bp1, bp2);
            o.put(" first  seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s - inc1 * (n_hom - 1)); o.put(','); o.num(s); o.put("] to");
                o.put(" ["); o.num(e); o.put(','); o.num(e + inc1 * (n_hom - 1)); o.put(']');
            }
            o.put('\n');
            s = frags[k].end2; e = frags[k + 1].start2;
            bp1 = inc2 * (s - s2.start); bp2 = inc2 * (e - s2.start);
            n_hom = outside_identity(s2, bp1, bp2);
            o.put(" second seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s - inc2 * (n_hom - 1)); o.put(','); o.num(s); o.put("] to");
                o.put(" ["); o.num(e); o.put(','); o.num(e + inc2 * (n_hom - 1)); o.put(']');
            }
            o.put('\n');
        }

        o.put("Identity inside breakpoints: \n");
        for (int k = 0; k + 1 < n_frg; k++) {
            int s, e;
            if (!excised_range(flag, s1.reverse, frags[k], frags[k + 1], s, e)) continue;
            int bp1 = inc1 * (s - s1.start), bp2 = inc1 * (e - s1.start);
            int n_hom = inside_identity(s1, bp1, bp2);
            o.put(" first  seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s); o.put(','); o.num(s + inc1 * (n_hom - 1)); o.put("] to");
                o.put(" ["); o.num(e - inc1 * (n_hom - 1)); o.put(','); o.num(e); o.put(']');
            }
            o.put('\n');
            s = frags[k].end2 + inc2; e = frags[k + 1].start2 - inc2;
            bp1 = inc2 * (s - s2.start); bp2 = inc2 * (e - s2.start);
            n_hom = inside_identity(s2, bp1, bp2);
            o.put(" second seq => "); o.num(n_hom, 9); o.put(" nucs");
            if (n_hom > 0) {
                o.put(" ["); o.num(s); o.put(','); o.num(s + inc2 * (n_hom - 1)); o.put("] to");
                o.put(" ["); o.num(e - inc2 * (n_hom - 1)); o.put(','); o.num(e); o.put(']');
            }
            o.put('\n');
        }
    }

    for (int k = 0; k < n_frg; k++) {
        if (k) o.put("\nEXCISED REGION\n");
        print_frag(o, frags[k]);
    }
    if (o.n < cap) buf[o.n] = '\0'; else if (cap) buf[cap - 1] = '\0';
    return o.n;
}

// AGEaligner::SetAlignResult of the AsmvarDetect flavour (src/AsmvarDetect/AGEaligner.cpp:179-332): the same quantities
// printAlignment() prints, as numbers.  Text (ali1 / ali2 / mapInfo per fragment) goes to `text`.
size_t describe_alignment(const age_seqview *v1, const age_seqview *v2, const age_result *res, age_align_info *info, char *text, size_t cap)
{
    memset(info, 0, sizeof(*info));
    info->score = res->score;                                   // score(): the aux score only if strictly greater
    info->strand = '.';
    const int flag = res->flag;
    SeqHdr s1{v1->seq, v1->len, v1->start, v1->reverse != 0};
    SeqHdr s2{v2->seq, v2->len, v2->start, v2->reverse != 0};
    std::string rc;
    SeqHdr s1rc = s1;
    if (flag & (AGE_INVL_FLAG | AGE_INVR_FLAG)) {
        rc.resize((size_t)s1.len);
        for (int i = 0; i < s1.len; i++) rc[(size_t)i] = complement(s1.seq[s1.len - 1 - i]);
        s1rc.seq = rc.data();
        s1rc.start = s1.reverse ? s1.start - (s1.len - 1) : s1.start + (s1.len - 1);
        s1rc.reverse = !s1.reverse;
    }
    std::vector<Frag> frags;
    build_frags(s1, s1rc, s2, flag, res->left, res->n_left, res->right, res->n_right, frags);
    const int n_frg = (int)frags.size();
    info->n_frag = n_frg;
    Out o{text, cap, 0};
    int n_ali = 0, n_id = 0;
    for (int k = 0; k < n_frg; k++) {
        const Frag &f = frags[k];
        age_fragment &g = info->frag[k];
        g.start1 = f.start1; g.end1 = f.end1; g.start2 = f.start2; g.end2 = f.end2;
        g.text_off = (int32_t)o.n; g.text_len = (int32_t)f.a1.size();
        for (size_t i = 0; i < f.a1.size(); i++) {              // AliFragment::countAligned
            g.n_aligned++;
            if (same_nuc(f.a1[i], f.a2[i])) g.n_identic++;
            if (is_gap(f.a1[i]) || is_gap(f.a2[i])) g.n_gap++;
        }
        n_ali += g.n_aligned; n_id += g.n_identic;
        o.put(f.a1.data(), f.a1.size()); o.put('\0');
        o.put(f.a2.data(), f.a2.size()); o.put('\0');
        for (size_t i = 0; i < f.a1.size(); i++)                // AliFragment::mapInfo (AGEaligner.h:318-334)
            o.put(same_nuc(f.a1[i], f.a2[i]) ? '|' : (!is_gap(f.a1[i]) && !is_gap(f.a2[i])) ? '.' : ' ');
        o.put('\0');
    }
    if (n_ali > 0) {
        info->identity[info->n_identity][0] = n_id; info->identity[info->n_identity][1] = (int)(100. * n_id / n_ali + 0.5);
        info->n_identity++;
        info->strand = info->frag[0].start2 > info->frag[0].end2 ? '-' : '+';
    }
    if (n_frg > 1) {
        for (int k = 0; k < n_frg; k++) {
            info->identity[info->n_identity][0] = info->frag[k].n_identic;
            info->identity[info->n_identity][1] = (int)(100.0 * info->frag[k].n_identic / info->frag[k].n_aligned + 0.5);
            info->n_identity++;
        }
        info->is_alternative_align = res->n_bp > 1;
        const int inc1 = s1.reverse ? -1 : 1, inc2 = s2.reverse ? -1 : 1;
        struct CI { bool any = false; int lo = 0, hi = 0; void add(int a, int b) { if (!any) { lo = a; hi = b; any = true; } else { if (lo > a) lo = a; if (hi < b) hi = b; } } };
        CI cs1, ce1, cs2, ce2;                                  // Boundary(): min of the firsts, max of the seconds
        const Frag &f = frags[0], &nx = frags[1];
        int s, e, left = 0, right = 0, h;
        if (excised_range(flag, s1.reverse, f, nx, s, e, 0, 1)) {                       // at the breakpoints
            h = at_identity(s1, inc1 * (s - s1.start), inc1 * (e - s1.start), left, right);
            info->homo_run_atbp1 = h;
            if (h > 0) { cs1.add(s + inc1 * left, s + inc1 * (right - 1)); ce1.add(e + inc1 * left, e + inc1 * (right - 1)); }
            s = f.end2 + inc2; e = nx.start2; left = right = 0;
            h = at_identity(s2, inc2 * (s - s2.start), inc2 * (e - s2.start), left, right);
            info->homo_run_atbp2 = h;
            if (h > 0) { cs2.add(s + inc2 * left, s + inc2 * (right - 1)); ce2.add(e + inc2 * left, e + inc2 * (right - 1)); }
        }
        if (excised_range(flag, s1.reverse, f, nx, s, e, -1, 1)) {                      // outside
            h = outside_identity(s1, inc1 * (s - s1.start), inc1 * (e - s1.start));
            info->homo_run_outbp1 = h;
            if (h > 0) { cs1.add(s - inc1 * (h - 1), s); ce1.add(e, e + inc1 * (h - 1)); }
            s = f.end2; e = nx.start2;
            h = outside_identity(s2, inc2 * (s - s2.start), inc2 * (e - s2.start));
            info->homo_run_outbp2 = h;
            if (h > 0) { cs2.add(s - inc2 * (h - 1), s); ce2.add(e, e + inc2 * (h - 1)); }
        }
        if (excised_range(flag, s1.reverse, f, nx, s, e)) {                             // inside
            h = inside_identity(s1, inc1 * (s - s1.start), inc1 * (e - s1.start));
            info->homo_run_inbp1 = h;
            if (h > 0) { cs1.add(s, s + inc1 * (h - 1)); ce1.add(e - inc1 * (h - 1), e); }
            s = f.end2 + inc2; e = nx.start2 - inc2;
            h = inside_identity(s2, inc2 * (s - s2.start), inc2 * (e - s2.start));
            info->homo_run_inbp2 = h;
            if (h > 0) { cs2.add(s, s + inc2 * (h - 1)); ce2.add(e - inc2 * (h - 1), e); }
        }
        info->ci_start1[0] = cs1.lo; info->ci_start1[1] = cs1.hi; info->ci_end1[0] = ce1.lo; info->ci_end1[1] = ce1.hi;
        info->ci_start2[0] = cs2.lo; info->ci_start2[1] = cs2.hi; info->ci_end2[0] = ce2.lo; info->ci_end2[1] = ce2.hi;
    }
    return o.n;
}

// ---- variant re-calling from the excised region -----------------------------------------------------
// Semantics: AgeAlignment::VarReCall / CallVarInExcise (src/AsmvarDetect/VarUnit.cpp:339-466) on top of the AlignResult
// that SetAlignResult fills (src/AsmvarDetect/AGEaligner.cpp:179-332).  Everything is derived from the block lists and
// the numbers the device left in age_result::counts; no gapped string is built.
int recall_variant(const age_seqview *v1, const age_seqview *v2, const age_result *res, age_variant *out)
{
    memset(out, 0, sizeof(*out));
    out->strand = '.';
    if (!res || res->status != AGE_OK || !res->detail) return AGE_ERR_ARG;
    const int flag = res->flag;
    const age_counts &cn = res->counts;
    const int step1 = v1->reverse ? -1 : 1, step2 = v2->reverse ? -1 : 1;
    // coordinates of the fragments (findAlignment, AGEaligner.cpp:686-700): the right fragment of INVL and the left one of
    // INVR live on the reverse complement of the first sequence, whose header starts at the other end and runs the other way
    const age_block *bl[2] = {res->left, res->right};
    const int nb[2] = {res->n_left, res->n_right};
    Frag fr[2];
    int side[2], n_frag = 0;
    for (int k = 0; k < 2; k++) {
        if (nb[k] <= 0 || !bl[k] || cn.n_aligned[k] <= 0) continue;
        const bool rc = ((flag & AGE_INVL_FLAG) && k == 1) || ((flag & AGE_INVR_FLAG) && k == 0);
        const int origin = rc ? v1->start + step1 * (v1->len - 1) : v1->start, dir = rc ? -step1 : step1;
        const age_block &first = bl[k][0], &last = bl[k][nb[k] - 1];
        Frag &f = fr[n_frag];
        f.start1 = origin + dir * (first.start1 - 1);
        f.end1 = origin + dir * (last.start1 + last.length - 2);
        f.start2 = v2->start + step2 * (first.start2 - 1);
        f.end2 = v2->start + step2 * (last.start2 + last.length - 2);
        side[n_frag++] = k;
    }
    out->n_frag = n_frag;
    long ali = 0, idn = 0;
    for (int i = 0; i < n_frag; i++) {
        const Frag &f = fr[i];
        out->frag[i][0] = f.start1; out->frag[i][1] = f.end1; out->frag[i][2] = f.start2; out->frag[i][3] = f.end2;
        ali += cn.n_aligned[side[i]]; idn += cn.n_identic[side[i]];
    }
    auto pct = [](long n, long of) { return (int)(100.0 * n / of + 0.5); };
    if (ali > 0) {
        out->identity[0][0] = (int)idn; out->identity[0][1] = pct(idn, ali); out->n_identity = 1;
        out->strand = fr[0].start2 > fr[0].end2 ? '-' : '+';
    }
    if (n_frag < 2) return 0;
    for (int i = 0; i < 2; i++) {
        out->identity[1 + i][0] = cn.n_identic[side[i]];
        out->identity[1 + i][1] = pct(cn.n_identic[side[i]], cn.n_aligned[side[i]]);
    }
    out->n_identity = 3;
    out->good_align = out->identity[0][1] > 98 && out->identity[1][0] > 30 && out->identity[1][1] > 95 &&
                      out->identity[2][0] > 30 && out->identity[2][1] > 95;

    // the span the breakpoints can slide over: union of what the three runs allow (SetAlignResult's Boundary())
    struct Span { bool any = false; int lo = 0, hi = 0; void add(int a, int b) { if (!any) { lo = a; hi = b; any = true; } else { lo = std::min(lo, a); hi = std::max(hi, b); } } };
    Span s1, e1, s2, e2;
    const Frag &f = fr[0], &nx = fr[1];
    int s, e;
    if (excised_range(flag, v1->reverse != 0, f, nx, s, e, 0, 1)) {                     // at the breakpoints
        int h = cn.homo_at1[0];
        if (h > 0) { s1.add(s + step1 * cn.homo_at1[1], s + step1 * (cn.homo_at1[2] - 1)); e1.add(e + step1 * cn.homo_at1[1], e + step1 * (cn.homo_at1[2] - 1)); }
        s = f.end2 + step2; e = nx.start2; h = cn.homo_at2[0];
        if (h > 0) { s2.add(s + step2 * cn.homo_at2[1], s + step2 * (cn.homo_at2[2] - 1)); e2.add(e + step2 * cn.homo_at2[1], e + step2 * (cn.homo_at2[2] - 1)); }
    }
    if (excised_range(flag, v1->reverse != 0, f, nx, s, e, -1, 1)) {                    // outside
        int h = cn.homo_out1;
        if (h > 0) { s1.add(s - step1 * (h - 1), s); e1.add(e, e + step1 * (h - 1)); }
        s = f.end2; e = nx.start2; h = cn.homo_out2;
        if (h > 0) { s2.add(s - step2 * (h - 1), s); e2.add(e, e + step2 * (h - 1)); }
    }
    if (excised_range(flag, v1->reverse != 0, f, nx, s, e)) {                           // inside
        int h = cn.homo_in1;
        if (h > 0) { s1.add(s, s + step1 * (h - 1)); e1.add(e - step1 * (h - 1), e); }
        s = f.end2 + step2; e = nx.start2 - step2; h = cn.homo_in2;
        if (h > 0) { s2.add(s, s + step2 * (h - 1)); e2.add(e - step2 * (h - 1), e); }
    }
    out->ci_start1[0] = s1.lo; out->ci_start1[1] = s1.hi; out->ci_end1[0] = e1.lo; out->ci_end1[1] = e1.hi;
    out->ci_start2[0] = s2.lo; out->ci_start2[1] = s2.hi; out->ci_end2[0] = e2.lo; out->ci_end2[1] = e2.hi;
    out->homo_run = cn.homo_at1[0];

    // The variant in the excised region.  What lies between the fragments on the target (tlen bases) and on the query
    // (qlen bases) decides the type; events that change the length are reported with the base in front of them (so that a
    // VCF line needs no extra reference base), same-length substitutions without it.  On the reverse strand the query
    // coordinates run downwards, so the query start is taken from the right fragment.
    out->has_variant = 1;
    const int tlen = nx.start1 - f.end1 - 1, qlen = std::abs(nx.start2 - f.end2) - 1;
    out->tlen = tlen; out->qlen = qlen;
    const bool forward = out->strand == '+';
    if (tlen > 0 && qlen == 0)      out->type = AGE_VAR_DEL;
    else if (qlen > 0 && tlen == 0) out->type = AGE_VAR_INS;
    else if (tlen == qlen)          out->type = tlen == 1 ? AGE_VAR_SNP : AGE_VAR_MNP;
    else                            out->type = AGE_VAR_REPLACEMENT;
    const int lead = (out->type == AGE_VAR_SNP || out->type == AGE_VAR_MNP) ? 1 : 0;   // no anchor base: start one further, end one earlier
    out->target_start = f.end1 + lead;
    out->target_end = out->target_start + tlen - lead;
    if (out->type == AGE_VAR_DEL) out->query_start = out->query_end = f.end2;
    else {
        out->query_start = forward ? f.end2 + lead : nx.start2 + 1;
        out->query_end = out->query_start + qlen - lead;
    }
    out->cipos[0] = out->ci_start1[0] > 0 ? out->ci_start1[0] - (int)out->target_start : 0;
    out->cipos[1] = out->ci_start1[1] > 0 ? out->ci_start1[1] - (int)out->target_start : 0;
    out->ciend[0] = out->ci_end1[0] > 0 ? out->ci_end1[0] - (int)out->target_end : 0;
    out->ciend[1] = out->ci_end1[1] > 0 ? out->ci_end1[1] - (int)out->target_end : 0;
    return 0;
}

} // namespace age

extern "C" int age_recall_variant(const age_seqview *s1, const age_seqview *s2, const age_result *res, age_variant *out)
{
    if (!s1 || !s2 || !out) return AGE_ERR_ARG;
    return age::recall_variant(s1, s2, res, out);
}

extern "C" const char *age_variant_type_name(int type)
{
    static const char *names[] = {"", "DEL", "INS", "SNP", "MNP", "REPLACEMENT"};
    return (type >= 0 && type <= AGE_VAR_REPLACEMENT) ? names[type] : "";
}

extern "C" size_t age_format_vcf_age(const age_variant *v, int good_realign, char *buf, size_t cap)
{
    char tmp[160];
    int n;
    if (!v || v->n_identity < 3) n = snprintf(tmp, sizeof tmp, ".");
    else n = snprintf(tmp, sizeof tmp, "%c,%c,%d,%d,%d,%d,%d,%d", good_realign ? 'T' : 'F', (char)v->strand, v->identity[0][0], v->identity[0][1],
                      v->identity[1][0], v->identity[1][1], v->identity[2][0], v->identity[2][1]);
    if (buf && cap) { const size_t m = std::min((size_t)n, cap - 1); memcpy(buf, tmp, m); buf[m] = '\0'; }
    return (size_t)n;
}

extern "C" size_t age_format_alignment(const age_seqview *s1, const age_seqview *s2, const age_job *job,
                                       const age_result *res, char *buf, size_t cap)
{
    if (!s1 || !s2 || !job || !res) return 0;
    return age::format_alignment(s1, s2, job, res, buf, cap);
}

extern "C" size_t age_describe_alignment(const age_seqview *s1, const age_seqview *s2, const age_result *res,
                                         age_align_info *info, char *text, size_t cap)
{
    if (!s1 || !s2 || !res || !info) return 0;
    return age::describe_alignment(s1, s2, res, info, text, cap);
}

extern "C" void age_revcom(const char *in, int32_t n, char *out)
{
    for (int32_t i = 0; i < n; i++) out[i] = age::complement(in[n - 1 - i]);
}

extern "C" int age_subst_score(char a, char b, int match, int mismatch)
{   // Scorer.hh:24-39
    if (age::complement(a) == a || age::complement(b) == b) return 0;
    return age::same_nuc(a, b) ? match : mismatch;
}
