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

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace age_b200 {

class Sequence {                                  // Sequence.hh:20-215 (the subset the aligner and its callers use)
public:
    Sequence(const std::string &seq, const std::string &name = "", int start = 1, bool rev = false)
        : _seq(seq), _name(name), _start(start), _reverse(rev) {}
    std::string &sequence() { return _seq; }
    const std::string &sequence() const { return _seq; }
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

// AsmvarDetect's structured result (src/AsmvarDetect/AGEaligner.h:70-150): same member names and defaults
class MapData {
public:
    MapData() : _start(0), _end(0) {}
    std::string _id;
    int _start, _end;
    std::string _sequence;
};

class AlignResult {
public:
    AlignResult() : _score(-1), _strand('.'), _is_alternative_align(false), _ci_start1(0, 0), _ci_end1(0, 0), _ci_start2(0, 0), _ci_end2(0, 0),
                    _homo_run_atbp1(0), _homo_run_atbp2(0), _homo_run_inbp1(0), _homo_run_inbp2(0), _homo_run_outbp1(0), _homo_run_outbp2(0) {}
    std::string _id1, _id2;
    int _score;
    char _strand;
    bool _is_alternative_align;
    std::vector<std::pair<int, int> > _identity;              // <identical bases, identity %>: average, left, right
    std::pair<int, int> _ci_start1, _ci_end1, _ci_start2, _ci_end2;
    int _homo_run_atbp1, _homo_run_atbp2, _homo_run_inbp1, _homo_run_inbp2, _homo_run_outbp1, _homo_run_outbp2;
    std::vector<std::pair<MapData, MapData> > _map;           // first -> seq1, second -> seq2
    std::vector<std::string> _map_info;
};

class AGEaligner {                                // AGEaligner.hh:69-126, src/AsmvarDetect/AGEaligner.h:152-227
public:
    static const int INDEL_FLAG = AGE_INDEL_FLAG, TDUPLICATION_FLAG = AGE_TDUPLICATION_FLAG, INVERSION_FLAG = AGE_INVERSION_FLAG,
                     INVR_FLAG = AGE_INVR_FLAG, INVL_FLAG = AGE_INVL_FLAG;

    AGEaligner(Sequence &s1, Sequence &s2) : _s1(s1), _s2(s2), _aligned(false), _is_set_align_result(false) {}

    // AGEaligner.cpp:65-147.  false = nothing to align (empty sequence, no mode).  Jobs outside the 16-bit score range
    // of the GPU kernels throw: the reference would wrap silently (AGEaligner.hh:86).
    bool align(Scorer &scr, int flag) {
        const age_job job = make_job(scr, flag);
        age_result res;
        if (age_align_batch(context(), &job, 1, &res) != 0) throw std::runtime_error(age_last_error(context()));
        return adopt(job, res);
    }
    // The two halves of align() for callers that batch: make_job() describes this aligner's work, adopt() takes the
    // result of that job out of an age_align_batch / age_collect result array (copies the block lists).
    age_job make_job(Scorer &scr, int flag) const {
        age_job j;
        j.seq1 = _s1.sequence().data(); j.len1 = (int32_t)_s1.sequence().size();
        j.seq2 = _s2.sequence().data(); j.len2 = (int32_t)_s2.sequence().size();
        j.flag = flag;
        j.match = scr.getMatch(); j.mismatch = scr.getMismatch(); j.gap_open = scr.getGapOpen(); j.gap_extend = scr.getGapExtend();
        j.partner = -1;
        return j;
    }
    bool adopt(const age_job &job, const age_result &res) {
        _job = job; _res = res;
        _is_set_align_result = false;
        if (_res.status < 0) throw std::range_error("AGEaligner: job outside the supported score range");
        _aligned = _res.status == AGE_OK;
        // the block arrays belong to the context until its next batch: keep copies
        _blocks.clear();
        if (_res.n_left) _blocks.insert(_blocks.end(), _res.left, _res.left + _res.n_left);
        if (_res.n_right) _blocks.insert(_blocks.end(), _res.right, _res.right + _res.n_right);
        if (_res.n_alt_left) _blocks.insert(_blocks.end(), _res.alt_left, _res.alt_left + _res.n_alt_left);
        if (_res.n_alt_right) _blocks.insert(_blocks.end(), _res.alt_right, _res.alt_right + _res.n_alt_right);
        return _aligned;
    }
    int score() const { return _aligned ? _res.score : 0; }                  // AGEaligner.cpp:149-154
    void printAlignment(std::ostream &os = std::cout) {                        // AGEaligner.cpp:179-380
        if (!_aligned) return;
        const age_result r = owned_result();
        const age_seqview v1 = view1(), v2 = view2();
        std::string buf(8192 + 8 * (_s1.sequence().size() + _s2.sequence().size()), '\0');
        size_t n = age_format_alignment(&v1, &v2, &_job, &r, &buf[0], buf.size());
        if (n >= buf.size()) { buf.resize(n + 1); n = age_format_alignment(&v1, &v2, &_job, &r, &buf[0], buf.size()); }
        os.write(buf.data(), (std::streamsize)n);
    }
    const age_result &result() const { return _res; }                         // breakpoints and blocks, without re-parsing text
    // The variant in the excised region and the figures around it (age_recall_variant: AgeAlignment::VarReCall /
    // CallVarInExcise, src/AsmvarDetect/VarUnit.cpp:339-466), from the block lists and the counts the device computed.
    // false where the reference's AlignResult would be empty: nothing aligned, or the auxiliary INVR aligner printed.
    bool recall(age_variant &v) const {
        if (!_aligned || _res.used_aux) return false;
        const age_result r = owned_result();
        const age_seqview v1 = view1(), v2 = view2();
        return age_recall_variant(&v1, &v2, &r, &v) == 0;
    }
    const std::string &name1() const { return _s1.name(); }
    const std::string &name2() const { return _s2.name(); }

    // src/AsmvarDetect/AGEaligner.cpp:179-332.  As there: when the auxiliary INVR aligner of -inv scored higher, its
    // record is printed and the AlignResult keeps its defaults.
    void SetAlignResult(std::ostream &os = std::cout) {
        _is_set_align_result = true;
        _align_result = AlignResult();
        if (!_aligned) return;
        if (_res.used_aux) { printAlignment(os); return; }
        const age_result r = owned_result();
        const age_seqview v1 = view1(), v2 = view2();
        age_align_info info;
        std::string text(64 + 6 * (_s1.sequence().size() + _s2.sequence().size()), '\0');
        size_t n = age_describe_alignment(&v1, &v2, &r, &info, &text[0], text.size());
        if (n > text.size()) { text.resize(n); age_describe_alignment(&v1, &v2, &r, &info, &text[0], text.size()); }
        AlignResult &a = _align_result;
        a._score = info.score;
        a._id1 = _s1.name(); a._id2 = _s2.name();
        a._strand = (char)info.strand;
        a._is_alternative_align = info.is_alternative_align != 0;
        for (int i = 0; i < info.n_identity; ++i) a._identity.push_back(std::make_pair((int)info.identity[i][0], (int)info.identity[i][1]));
        a._ci_start1 = std::make_pair((int)info.ci_start1[0], (int)info.ci_start1[1]); a._ci_end1 = std::make_pair((int)info.ci_end1[0], (int)info.ci_end1[1]);
        a._ci_start2 = std::make_pair((int)info.ci_start2[0], (int)info.ci_start2[1]); a._ci_end2 = std::make_pair((int)info.ci_end2[0], (int)info.ci_end2[1]);
        a._homo_run_atbp1 = info.homo_run_atbp1; a._homo_run_atbp2 = info.homo_run_atbp2;
        a._homo_run_inbp1 = info.homo_run_inbp1; a._homo_run_inbp2 = info.homo_run_inbp2;
        a._homo_run_outbp1 = info.homo_run_outbp1; a._homo_run_outbp2 = info.homo_run_outbp2;
        for (int k = 0; k < info.n_frag; ++k) {
            const age_fragment &f = info.frag[k];
            MapData m1, m2;
            m1._id = a._id1; m2._id = a._id2;
            m1._start = f.start1; m1._end = f.end1; m2._start = f.start2; m2._end = f.end2;
            m1._sequence.assign(text, (size_t)f.text_off, (size_t)f.text_len);
            m2._sequence.assign(text, (size_t)f.text_off + f.text_len + 1, (size_t)f.text_len);
            a._map.push_back(std::make_pair(m1, m2));
            a._map_info.push_back(text.substr((size_t)f.text_off + 2 * ((size_t)f.text_len + 1), (size_t)f.text_len));
        }
    }
    AlignResult align_result() const {            // src/AsmvarDetect/AGEaligner.h:194-201
        if (!_is_set_align_result) {
            std::cerr << "[ERROR] You should Call SetAlignResult() first before ";
            std::cerr << "calling align_result() or the align_result would be empty\n";
            exit(1);
        }
        return _align_result;
    }

    static age_ctx *context() {                   // one context per process (the reference has no such state)
        static age_ctx *ctx = age_create(nullptr, 0);
        if (!ctx) throw std::runtime_error(age_last_error(nullptr));
        return ctx;
    }
private:
    age_result owned_result() const {             // _res with the block pointers redirected to our copies
        age_result r = _res;
        const age_block *b = _blocks.data();
        r.left = b; r.right = b + r.n_left; r.alt_left = r.right + r.n_right; r.alt_right = r.alt_left + r.n_alt_left;
        return r;
    }
    age_seqview view1() const { return age_seqview{_s1.sequence().data(), (int32_t)_s1.sequence().size(), _s1.name().c_str(), _s1.start(), _s1.reverse() ? 1 : 0}; }
    age_seqview view2() const { return age_seqview{_s2.sequence().data(), (int32_t)_s2.sequence().size(), _s2.name().c_str(), _s2.start(), _s2.reverse() ? 1 : 0}; }
    Sequence &_s1, &_s2;
    age_job _job;
    age_result _res;
    std::vector<age_block> _blocks;
    bool _aligned, _is_set_align_result;
    AlignResult _align_result;
};

} // namespace age_b200
#endif
