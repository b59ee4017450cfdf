// age_varunit.hh -- AsmvarDetect's entry to the realigner, on top of include/age_shim.hh.
//
// AsmvarDetect reaches the hot path through VarUnit::ReAlignAndReCallVar (src/AsmvarDetect/VarUnit.cpp:66-82), which
// wraps class AgeAlignment (VarUnit.h:121-160, VarUnit.cpp:215-470): extend the variant window by a flank, refuse
// windows whose matrices would need more than 10 GB, cut both windows out of the FASTA records, run one or two
// AGEaligners (-both), print `# type:strand:score:mismap:target:query` + the .age record, and re-call the variant
// from the excised region (DEL / INS / SNP / MNP / REPLACEMENT, good-alignment rule, confidence intervals).
// Same names, members and defaults here, in namespace age_b200; Variant::AGE_Realign (Variant.cpp:375-423) loops
// over variants one at a time -- ReAlignAndReCallVarBatch() below is that loop as ONE GPU batch.
#ifndef AGE_VARUNIT_HH
#define AGE_VARUNIT_HH

#include "age_shim.hh"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace age_b200 {

class Region {                                    // src/AsmvarDetect/Region.h:17-45
public:
    std::string id;
    long int start, end, regsize;
    std::string info;
    Region() : start(0), end(0), regsize(0) {}
    void SetRegSize() { regsize = (start + end > 0) ? labs(end - start) + 1 : 0; }
    bool isEmpty() const { return end == 0; }
};

class AgeOption {                                 // src/AsmvarDetect/AgeOption.h:9-32
public:
    int gapOpen, gapExtend, match, mismatch;
    int extendVarFlankSzie;                       // sic
    bool indel, inv, invl, invr, tdup;
    bool both, revcomTarget, revcomQuery;
    AgeOption() : gapOpen(-10), gapExtend(-1), match(1), mismatch(-1), extendVarFlankSzie(500), indel(true), inv(false), invl(false),
                  invr(false), tdup(false), both(true), revcomTarget(false), revcomQuery(false) {}
};

class VarUnit {                                   // src/AsmvarDetect/VarUnit.h:32-110, VarUnit.cpp:3-44
public:
    Region target, query;
    std::string tarSeq, qrySeq;
    char strand;
    std::string type;
    bool isHete;
    long score;                                   // LAST alignment score
    double mismap;                                // LAST mismap probability
    int homoRun;
    bool isSuccessAlign, isGoodReAlign;
    std::pair<int, int> cipos, ciend;
    std::vector<std::pair<int, int> > identity;
    Region exp_target;
    std::string exp_tarSeq;

    VarUnit() : tarSeq("."), qrySeq("."), strand('.'), type("."), isHete(false), score(0), mismap(1.0), homoRun(0), isSuccessAlign(false),
                isGoodReAlign(false), cipos(0, 0), ciend(0, 0) { target.id = "-"; query.id = "-"; }

    void PrintStd(std::ostream &os = std::cout) const {       // VarUnit.cpp:131-142
        if (tarSeq.empty() || qrySeq.empty()) { std::cerr << "tarSeq.empty() || qrySeq.empty()" << std::endl; exit(1); }
        os << "# " << type << ":" << strand << ":" << score << ":" << mismap << ":" << target.id << "-" << target.start << "-" << target.end
           << ":" << query.id << "-" << query.start << "-" << query.end << "\n";
    }
    // VarUnit.cpp:66-82; one single-candidate GPU batch.  Prefer ReAlignAndReCallVarBatch().
    std::vector<VarUnit> ReAlignAndReCallVar(std::string &targetSeq, std::string &querySeq, AgeOption opt, std::ostream &os = std::cout);
};

class AgeAlignment {                              // VarUnit.h:121-160
public:
    AgeAlignment() : isInit_(false), isalign_(false), isgoodAlign_(false), huge_(false) {}
    AgeAlignment(VarUnit &v, AgeOption opt) : isInit_(false), isalign_(false), isgoodAlign_(false), huge_(false) { Init(v, opt); }
    VarUnit vu() { return vu_; }
    void Init(VarUnit &v, AgeOption opt) { vu_ = v; rvu_ = v; para_ = opt; isInit_ = true; }

    // VarUnit.cpp:223-337 in one call
    bool Align(std::string &tarFa, std::string &qryFa, std::ostream &os = std::cout) {
        std::vector<age_job> jobs;
        if (!Prepare(tarFa, qryFa, jobs)) return false;
        std::vector<age_result> res(jobs.size());
        if (age_align_batch(AGEaligner::context(), jobs.data(), jobs.size(), res.data()) != 0) throw std::runtime_error(age_last_error(AGEaligner::context()));
        return Finish(jobs.data(), res.data(), os);
    }

    // First half of Align(): mode check, ExtendVU, the 10 GB guard, window extraction.  Appends this candidate's one
    // (or, with -both, two) jobs; false = not aligned (too large).
    bool Prepare(std::string &tarFa, std::string &qryFa, std::vector<age_job> &jobs) {
        if (!isInit_) { std::cerr << "[ERROR] You should init AgeAlignment before calling Align()\n"; exit(1); }
        flag_ = 0;
        if (para_.indel) flag_ |= AGEaligner::INDEL_FLAG;
        if (para_.inv) flag_ |= AGEaligner::INVERSION_FLAG;
        if (para_.invl) flag_ |= AGEaligner::INVL_FLAG;
        if (para_.invr) flag_ |= AGEaligner::INVR_FLAG;
        if (para_.tdup) flag_ |= AGEaligner::TDUPLICATION_FLAG;
        if (flag_ == 0) flag_ = AGEaligner::INDEL_FLAG;
        int n_modes = 0;
        for (int f : {AGE_INDEL_FLAG, AGE_TDUPLICATION_FLAG, AGE_INVR_FLAG, AGE_INVL_FLAG, AGE_INVERSION_FLAG}) if (flag_ & f) n_modes++;
        if (n_modes != 1) {
            std::cerr << "# [Error] In mode specification. ";
            if (n_modes > 1) std::cerr << "More than one mode is specified.\n";
            exit(1);
        }
        ExtendVU((long)tarFa.length(), (long)qryFa.length(), para_.extendVarFlankSzie);
        huge_ = IsHugeMemory(vu_.target.end - vu_.target.start, vu_.query.end - vu_.query.start);
        if (huge_) return false;                                    // "Do not AGE if the memory cost is bigger than 10G"
        Sequence tarSeq(tarFa, vu_.target.id, 1, false), qrySeq(qryFa, vu_.query.id, 1, false);
        tar_.reset(tarSeq.substr((int)vu_.target.start, (int)vu_.target.end));
        qry_.reset(qrySeq.substr((int)vu_.query.start, (int)vu_.query.end));
        if (!tar_ || !qry_) throw std::invalid_argument("AgeAlignment: empty variant window (start > end)");
        first_job_ = (int)jobs.size();
        aligner1_.reset(new AGEaligner(*tar_, *qry_));
        Scorer scr((short)para_.match, (short)para_.mismatch, (short)para_.gapOpen, (short)para_.gapExtend);
        jobs.push_back(aligner1_->make_job(scr, flag_));
        if (para_.both) {
            qryClone_.reset(qry_->clone());
            qryClone_->revcom();
            aligner2_.reset(new AGEaligner(*tar_, *qryClone_));
            jobs.push_back(aligner2_->make_job(scr, flag_));
            jobs[first_job_].partner = first_job_ + 1;              // only the strand that gets printed is traced back
            jobs[first_job_ + 1].partner = first_job_;
        }
        return true;
    }

    // Second half of Align(): choose the strand (VarUnit.cpp:283-299), SetAlignResult, print.  `jobs` / `res` are the
    // arrays of the batch Prepare() appended to.
    bool Finish(const age_job *jobs, const age_result *res, std::ostream &os = std::cout) {
        if (huge_) return false;
        isalign_ = true;
        const bool res1 = aligner1_->adopt(jobs[first_job_], res[first_job_]);
        AGEaligner *pick = nullptr;
        if (para_.both) {
            const bool res2 = aligner2_->adopt(jobs[first_job_ + 1], res[first_job_ + 1]);
            if (!res1 && !res2) { std::cerr << "No alignment made.\n"; isalign_ = false; }
            else pick = (aligner1_->score() >= aligner2_->score()) ? aligner1_.get() : aligner2_.get();
        } else {
            if (res1) pick = aligner1_.get();
            else { std::cerr << "No alignment made.\n"; isalign_ = false; }
        }
        picked_ = pick;
        if (pick) {
            pick->SetAlignResult(os);
            alignResult_ = pick->align_result();
            rvu_.PrintStd(os);
            pick->printAlignment(os);
            os << "\n";
        }
        return isalign_;
    }

    AlignResult AligResult() { return alignResult_; }
    bool isgoodAlign() { return isgoodAlign_; }
    int HomoRun() { return alignResult_._homo_run_atbp1; }
    std::pair<int, int> cipos() { return alignResult_._ci_start1; }
    std::pair<int, int> ciend() { return alignResult_._ci_end1; }

    // VarUnit.cpp:339-466 (VarReCall + CallVarInExcise): at most one variant, the one in the excised region between the two
    // fragments.  The classification, its coordinates, the confidence intervals and the good-alignment rule come from
    // age_recall_variant (computed from the block lists; the gapped strings are not consulted).
    std::vector<VarUnit> VarReCall() {
        std::vector<VarUnit> vus;
        age_variant v;
        if (!isalign_ || !picked_ || !picked_->recall(v)) return vus;
        isgoodAlign_ = v.good_align != 0;
        if (!v.has_variant) return vus;
        if (v.tlen < 0) std::cerr << "[WARNING] It'll cause bug below!\n";
        VarUnit var;
        var.type = age_variant_type_name(v.type);
        var.strand = (char)v.strand;
        var.score = vu_.score; var.mismap = vu_.mismap;           // of the alignment the candidate came from
        var.target.id = picked_->name1(); var.query.id = picked_->name2();
        var.target.start = (long)v.target_start; var.target.end = (long)v.target_end;
        var.query.start = (long)v.query_start; var.query.end = (long)v.query_end;
        var.isSuccessAlign = true;
        var.isGoodReAlign = isgoodAlign_ && var.mismap < 0.01;    // mis-mapping probability of the original alignment
        var.homoRun = v.homo_run;
        var.cipos = std::make_pair((int)v.cipos[0], (int)v.cipos[1]);
        var.ciend = std::make_pair((int)v.ciend[0], (int)v.ciend[1]);
        for (int i = 0; i < v.n_identity; ++i) var.identity.push_back(std::make_pair((int)v.identity[i][0], (int)v.identity[i][1]));
        vus.push_back(var);
        return vus;
    }

private:
    // VarUnit.cpp:585-603: widen both windows by the flank, inside the records
    void ExtendVU(long int tarFaSize, long int qryFaSize, int flank) {
        auto widen = [flank](Region &r, long int size) { r.start = std::max<long int>(1, r.start - flank); r.end = std::min<long int>(size, r.end + flank); };
        widen(vu_.target, tarFaSize);
        widen(vu_.query, qryFaSize);
    }
    // VarUnit.cpp:605-609: the reference's matrices take 5 bytes per cell; windows beyond ~10 GB (9.5 GB after its rounding) are skipped
    static bool IsHugeMemory(long int n, long int m) { return 5.0L * (long double)n * (long double)m > 9.5e9L; }

    VarUnit vu_, rvu_;
    AgeOption para_;
    AlignResult alignResult_;
    bool isInit_, isalign_, isgoodAlign_, huge_;
    int flag_ = 0, first_job_ = 0;
    AGEaligner *picked_ = nullptr;
    std::unique_ptr<Sequence> tar_, qry_, qryClone_;
    std::unique_ptr<AGEaligner> aligner1_, aligner2_;
};

inline std::vector<VarUnit> VarUnit::ReAlignAndReCallVar(std::string &targetSeq, std::string &querySeq, AgeOption opt, std::ostream &os)
{
    std::vector<VarUnit> vus;
    AgeAlignment alignment(*this, opt);
    if (alignment.Align(targetSeq, querySeq, os)) { isSuccessAlign = true; vus = alignment.VarReCall(); }
    else isSuccessAlign = false;
    return vus;
}

// The loop of Variant::AGE_Realign (Variant.cpp:375-423) as one GPU batch: vars[i] is realigned against
// *targets[i] / *queries[i] (whole FASTA records); returns what vars[i].ReAlignAndReCallVar() would, prints the
// same text in the same order.
inline std::vector<std::vector<VarUnit> > ReAlignAndReCallVarBatch(std::vector<VarUnit> &vars, const std::vector<std::string *> &targets,
                                                                   const std::vector<std::string *> &queries, AgeOption opt,
                                                                   std::ostream &os = std::cout)
{
    const size_t n = vars.size();
    std::vector<std::unique_ptr<AgeAlignment> > al(n);
    std::vector<char> prepared(n, 0);
    std::vector<age_job> jobs;
    for (size_t i = 0; i < n; ++i) {
        al[i].reset(new AgeAlignment(vars[i], opt));
        prepared[i] = al[i]->Prepare(*targets[i], *queries[i], jobs) ? 1 : 0;
    }
    std::vector<age_result> res(jobs.size() ? jobs.size() : 1);
    if (!jobs.empty() && age_align_batch(AGEaligner::context(), jobs.data(), jobs.size(), res.data()) != 0)
        throw std::runtime_error(age_last_error(AGEaligner::context()));
    std::vector<std::vector<VarUnit> > out(n);
    for (size_t i = 0; i < n; ++i) {
        vars[i].isSuccessAlign = prepared[i] && al[i]->Finish(jobs.data(), res.data(), os);
        if (vars[i].isSuccessAlign) out[i] = al[i]->VarReCall();
    }
    return out;
}

} // namespace age_b200
#endif
