// age_shim.hh -- the reference's in-process interface on top of the C ABI (include/age_b200.h).
//
// AsmVar's callers use three methods of class AGEaligner (src/AsmvarAGE/src/AGEaligner.hh:100-105):
//     AGEaligner(Sequence &s1, Sequence &s2);  bool align(Scorer &scr, int flag);  int score();  void printAlignment();
// with Sequence = {nucleotides, name, 1-based start, reverse flag} (Sequence.hh:22-31) and
// Scorer = {match, mismatch, gap open, gap extend} (Scorer.hh:17-47).  This header offers the same shapes in
// namespace age_b200 so that a caller written against the reference (age_align.cpp:244-283, VarUnit.cpp:264-299)
// needs a namespace change only.  One align() is one single-job GPU batch; callers that care about throughput
// collect jobs and call age_align_batch() directly.
#ifndef AGE_SHIM_HH
#define AGE_SHIM_HH

#include "age_b200.h"

#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace age_b200 {

class Sequence {                                  // Sequence.hh:20-215 (the subset the aligner and its callers use)
public:
    Sequence(const std::string &seq, const std::string &name = "", int start = 1, bool rev = false)
        : _seq(seq), _name(name), _start(start), _reverse(rev) {}
    std::string &sequence() { return _seq; }
    const std::string &name() const { return _name; }
    int start() const { return _start; }
    bool reverse() const { return _reverse; }
    Sequence *clone() const { return new Sequence(_seq, _name, _start, _reverse); }
    void revcom() {                               // Sequence.hh:190-198
        if (_reverse) _start -= (int)_seq.length() - 1; else _start += (int)_seq.length() - 1;
        _reverse = !_reverse;
        std::string r(_seq.size(), ' ');
        age_revcom(_seq.data(), (int32_t)_seq.size(), &r[0]);
        _seq.swap(r);
    }
    Sequence *substr(int start, int end) const {  // Sequence.hh:200-215, 1-based, <= 0 = open
        if (start > 0 && end > 0 && start > end) return nullptr;
        const int n = (int)_seq.length();
        if (start <= 0) start = 1;
        if (end <= 0 || end > n) end = n;
        int s = start, e = end;
        if (_reverse) { s = n - end + 1; e = n - start + 1; }
        return new Sequence(n > 0 ? _seq.substr(s - 1, e - s + 1) : std::string(), _name, _start + start - 1, _reverse);
    }
private:
    std::string _seq, _name;
    int _start;
    bool _reverse;
};

class Scorer {                                    // Scorer.hh:17-47
public:
    Scorer(short match, short mismatch, short gap_open, short gap_extend)
        : _match(match), _mismatch(mismatch), _gap_open(gap_open), _gap_extend(gap_extend) {}
    short getScore(char a, char b) const { return (short)age_subst_score(a, b, _match, _mismatch); }
    short getMatch() const { return _match; }
    short getMismatch() const { return _mismatch; }
    short getGapOpen() const { return _gap_open; }
    short getGapExtend() const { return _gap_extend; }
private:
    short _match, _mismatch, _gap_open, _gap_extend;
};

class AGEaligner {                                // AGEaligner.hh:69-126
public:
    static const int INDEL_FLAG = AGE_INDEL_FLAG, TDUPLICATION_FLAG = AGE_TDUPLICATION_FLAG, INVERSION_FLAG = AGE_INVERSION_FLAG,
                     INVR_FLAG = AGE_INVR_FLAG, INVL_FLAG = AGE_INVL_FLAG;

    AGEaligner(Sequence &s1, Sequence &s2) : _s1(s1), _s2(s2), _aligned(false) {}

    // AGEaligner.cpp:65-147.  false = nothing to align (empty sequence, no mode).  Jobs outside the 16-bit score range
    // of the GPU kernels throw: the reference would wrap silently (AGEaligner.hh:86).
    bool align(Scorer &scr, int flag) {
        _job.seq1 = _s1.sequence().data(); _job.len1 = (int32_t)_s1.sequence().size();
        _job.seq2 = _s2.sequence().data(); _job.len2 = (int32_t)_s2.sequence().size();
        _job.flag = flag;
        _job.match = scr.getMatch(); _job.mismatch = scr.getMismatch(); _job.gap_open = scr.getGapOpen(); _job.gap_extend = scr.getGapExtend();
        _job.partner = -1;
        if (age_align_batch(context(), &_job, 1, &_res) != 0) throw std::runtime_error(age_last_error(context()));
        if (_res.status < 0) throw std::range_error("AGEaligner: job outside the supported score range");
        _aligned = _res.status == AGE_OK;
        // the block arrays belong to the context until its next batch: keep copies
        _blocks.assign(_res.left, _res.left + _res.n_left);
        _blocks.insert(_blocks.end(), _res.right, _res.right + _res.n_right);
        _blocks.insert(_blocks.end(), _res.alt_left, _res.alt_left + _res.n_alt_left);
        _blocks.insert(_blocks.end(), _res.alt_right, _res.alt_right + _res.n_alt_right);
        return _aligned;
    }
    int score() const { return _aligned ? _res.score : 0; }                  // AGEaligner.cpp:149-154
    void printAlignment(std::ostream &os = std::cout) {                        // AGEaligner.cpp:179-380
        if (!_aligned) return;
        age_result r = _res;
        const age_block *b = _blocks.data();
        r.left = b; r.right = b + r.n_left; r.alt_left = r.right + r.n_right; r.alt_right = r.alt_left + r.n_alt_left;
        const age_seqview v1{_s1.sequence().data(), (int32_t)_s1.sequence().size(), _s1.name().c_str(), _s1.start(), _s1.reverse() ? 1 : 0};
        const age_seqview v2{_s2.sequence().data(), (int32_t)_s2.sequence().size(), _s2.name().c_str(), _s2.start(), _s2.reverse() ? 1 : 0};
        std::string buf(8192 + 8 * (_s1.sequence().size() + _s2.sequence().size()), '\0');
        size_t n = age_format_alignment(&v1, &v2, &_job, &r, &buf[0], buf.size());
        if (n >= buf.size()) { buf.resize(n + 1); n = age_format_alignment(&v1, &v2, &_job, &r, &buf[0], buf.size()); }
        os.write(buf.data(), (std::streamsize)n);
    }
    const age_result &result() const { return _res; }                         // breakpoints and blocks, without re-parsing text

    static age_ctx *context() {                   // one context per process (the reference has no such state)
        static age_ctx *ctx = age_create(nullptr, 0);
        if (!ctx) throw std::runtime_error(age_last_error(nullptr));
        return ctx;
    }
private:
    Sequence &_s1, &_s2;
    age_job _job;
    age_result _res;
    std::vector<age_block> _blocks;
    bool _aligned;
};

} // namespace age_b200
#endif
