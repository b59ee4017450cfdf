// AsmvarDetect's VarUnit::ReAlignAndReCallVar (src/AsmvarDetect/VarUnit.cpp:66-82) and the batched form of the
// Variant::AGE_Realign loop (Variant.cpp:375-423) against include/age_varunit.hh.
// stdin, one case per line (tab separated):
//   id tid tseq qid qseq tstart tend qstart qend type strand score mismap extend both flag match mismatch go ge
// stdout: one JSON object per case {id, success, text, vus:[...]}; with the argument `batch`, consecutive cases with the
// same options go through ReAlignAndReCallVarBatch() (one GPU batch) and `text` is reported once, on a final line
// {"all_text": ...}, because the records of a batch are printed back to back as in the reference's loop.
//   g++ -std=c++17 -Iinclude examples/shim_recall.cpp -Lasmvar_b200 -lage_b200 -Wl,-rpath,$PWD/asmvar_b200 -o shim_recall
#include "age_varunit.hh"

#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace age_b200;
using std::string;

static string esc(const string &s)
{
    string r;
    for (char c : s) {
        if (c == '"' || c == '\\') { r += '\\'; r += c; }
        else if (c == '\n') r += "\\n";
        else r += c;
    }
    return r;
}

struct Case { string id, tseq, qseq, optkey; VarUnit v; AgeOption opt; };

static void emit(const Case &c, const std::vector<VarUnit> &vus, const string *text)
{
    std::ostringstream o;
    o << "{\"id\":\"" << esc(c.id) << "\",\"success\":" << (c.v.isSuccessAlign ? 1 : 0);
    if (text) o << ",\"text\":\"" << esc(*text) << "\"";
    o << ",\"vus\":[";
    for (size_t i = 0; i < vus.size(); ++i) {
        const VarUnit &u = vus[i];
        o << (i ? "," : "") << "{\"type\":\"" << esc(u.type) << "\",\"strand\":\"" << u.strand << "\",\"tid\":\"" << esc(u.target.id)
          << "\",\"tstart\":" << u.target.start << ",\"tend\":" << u.target.end << ",\"qid\":\"" << esc(u.query.id) << "\",\"qstart\":"
          << u.query.start << ",\"qend\":" << u.query.end << ",\"score\":" << u.score << ",\"mismap\":" << u.mismap << ",\"homoRun\":"
          << u.homoRun << ",\"isSuccessAlign\":" << (u.isSuccessAlign ? 1 : 0) << ",\"isGoodReAlign\":" << (u.isGoodReAlign ? 1 : 0)
          << ",\"cipos\":[" << u.cipos.first << "," << u.cipos.second << "],\"ciend\":[" << u.ciend.first << "," << u.ciend.second << "],\"identity\":[";
        for (size_t k = 0; k < u.identity.size(); ++k) o << (k ? "," : "") << "[" << u.identity[k].first << "," << u.identity[k].second << "]";
        o << "]}";
    }
    o << "]}";
    std::cout << o.str() << "\n";
}

int main(int argc, char **argv)
{
    const bool batch = argc > 1 && strcmp(argv[1], "batch") == 0;
    std::vector<Case> cases;
    string line;
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::vector<string> f;
        { std::stringstream ss(line); string t; while (std::getline(ss, t, '\t')) f.push_back(t); }
        if (f.size() < 20) { std::cerr << "bad case line\n"; return 2; }
        Case c;
        c.id = f[0]; c.tseq = f[2]; c.qseq = f[4];
        c.v.target.id = f[1]; c.v.query.id = f[3];
        c.v.target.start = atol(f[5].c_str()); c.v.target.end = atol(f[6].c_str());
        c.v.query.start = atol(f[7].c_str()); c.v.query.end = atol(f[8].c_str());
        c.v.type = f[9]; c.v.strand = f[10][0]; c.v.score = atol(f[11].c_str()); c.v.mismap = atof(f[12].c_str());
        c.opt.extendVarFlankSzie = atoi(f[13].c_str());
        c.opt.both = atoi(f[14].c_str()) != 0;
        const int flag = atoi(f[15].c_str());
        c.opt.indel = flag & 0x01; c.opt.tdup = flag & 0x02; c.opt.inv = flag & 0x04; c.opt.invr = flag & 0x08; c.opt.invl = flag & 0x10;
        c.opt.match = atoi(f[16].c_str()); c.opt.mismatch = atoi(f[17].c_str()); c.opt.gapOpen = atoi(f[18].c_str()); c.opt.gapExtend = atoi(f[19].c_str());
        for (int k = 13; k < 20; ++k) c.optkey += f[k] + ",";
        cases.push_back(c);
    }
    if (!batch) {
        for (Case &c : cases) {
            std::ostringstream text;
            std::vector<VarUnit> vus = c.v.ReAlignAndReCallVar(c.tseq, c.qseq, c.opt, text);
            const string t = text.str();
            emit(c, vus, &t);
        }
        return 0;
    }
    std::ostringstream all;
    for (size_t i = 0; i < cases.size();) {
        size_t j = i;
        while (j < cases.size() && cases[j].optkey == cases[i].optkey) ++j;
        std::vector<VarUnit> vars;
        std::vector<string *> tg, qy;
        for (size_t k = i; k < j; ++k) { vars.push_back(cases[k].v); tg.push_back(&cases[k].tseq); qy.push_back(&cases[k].qseq); }
        std::vector<std::vector<VarUnit> > out = ReAlignAndReCallVarBatch(vars, tg, qy, cases[i].opt, all);
        for (size_t k = i; k < j; ++k) { cases[k].v = vars[k - i]; emit(cases[k], out[k - i], nullptr); }
        i = j;
    }
    std::cout << "{\"all_text\":\"" << esc(all.str()) << "\"}\n";
    return 0;
}
