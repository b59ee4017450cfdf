// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Driver around the UNMODIFIED AsmvarDetect sources (src/AsmvarDetect/{VarUnit,AGEaligner,Region}.cpp, compiled where
// they lie by oracle/build_ref.sh): VarUnit::ReAlignAndReCallVar (VarUnit.cpp:66-82) on cases read from stdin; dumps
// what it prints and the VarUnits it returns, one JSON object per line -> tests/golden/recall_cases.json.
//
// stdin, one case per line (tab separated):
//   id tid tseq qid qseq tstart tend qstart qend type strand score mismap extend both flag match mismatch go ge
#include "VarUnit.h"

#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

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

int main()
{
    string line;
    while (getline(cin, line)) {
        if (line.empty()) continue;
        vector<string> f;
        { stringstream ss(line); string t; while (getline(ss, t, '\t')) f.push_back(t); }
        if (f.size() < 20) { cerr << "bad case line" << endl; return 2; }
        VarUnit v;
        v.target.id = f[1]; v.query.id = f[3];
        v.target.start = atol(f[5].c_str()); v.target.end = atol(f[6].c_str());
        v.query.start = atol(f[7].c_str()); v.query.end = atol(f[8].c_str());
        v.type = f[9]; v.strand = f[10][0]; v.score = atol(f[11].c_str()); v.mismap = atof(f[12].c_str());
        AgeOption opt;
        opt.extendVarFlankSzie = atoi(f[13].c_str());
        opt.both = atoi(f[14].c_str()) != 0;
        const int flag = atoi(f[15].c_str());
        opt.indel = flag & 0x01; opt.tdup = flag & 0x02; opt.inv = flag & 0x04; opt.invr = flag & 0x08; opt.invl = flag & 0x10;
        opt.match = atoi(f[16].c_str()); opt.mismatch = atoi(f[17].c_str()); opt.gapOpen = atoi(f[18].c_str()); opt.gapExtend = atoi(f[19].c_str());
        streambuf *keep = cout.rdbuf();
        ostringstream text;
        cout.rdbuf(text.rdbuf());
        vector<VarUnit> vus = v.ReAlignAndReCallVar(f[2], f[4], opt);
        cout.rdbuf(keep);
        ostringstream o;
        o << "{\"id\":\"" << esc(f[0]) << "\",\"success\":" << (v.isSuccessAlign ? 1 : 0) << ",\"text\":\"" << esc(text.str()) << "\",\"vus\":[";
        for (size_t i = 0; i < vus.size(); ++i) {
            VarUnit &u = vus[i];
            o << (i ? "," : "") << "{\"type\":\"" << esc(u.type) << "\",\"strand\":\"" << u.strand << "\",\"tid\":\"" << esc(u.target.id)
              << "\",\"tstart\":" << u.target.start << ",\"tend\":" << u.target.end << ",\"qid\":\"" << esc(u.query.id) << "\",\"qstart\":"
              << u.query.start << ",\"qend\":" << u.query.end << ",\"score\":" << u.score << ",\"mismap\":" << u.mismap << ",\"homoRun\":"
              << u.homoRun << ",\"isSuccessAlign\":" << (u.isSuccessAlign ? 1 : 0) << ",\"isGoodReAlign\":" << (u.isGoodReAlign ? 1 : 0)
              << ",\"cipos\":[" << u.cipos.first << "," << u.cipos.second << "],\"ciend\":[" << u.ciend.first << "," << u.ciend.second << "],\"identity\":[";
            for (size_t k = 0; k < u.identity.size(); ++k) o << (k ? "," : "") << "[" << u.identity[k].first << "," << u.identity[k].second << "]";
            o << "]}";
        }
        o << "]}";
        cout << o.str() << "\n";
    }
    return 0;
}
