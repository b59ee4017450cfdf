// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Driver around the UNMODIFIED AsmvarDetect flavour of the reference aligner (src/AsmvarDetect/AGEaligner.{h,cpp},
// compiled where it lies by oracle/build_ref.sh): runs AGEaligner::align() + SetAlignResult() on cases read from stdin
// and dumps the AlignResult (src/AsmvarDetect/AGEaligner.h:91-150) as one JSON object per line.  The dumps are the
// golden vectors of include/age_shim.hh's SetAlignResult()/align_result() (tests/golden/align_result_cases.json).
//
// stdin, one case per line (tab separated):
//   id  name1  start1  reverse1  seq1  name2  start2  reverse2  seq2  flag  match  mismatch  go  ge
#include "AGEaligner.h"
#include "Sequence.h"
#include "Scorer.h"

#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

static string esc(const string &s)
{
    string r;
    for (char c : s) { if (c == '"' || c == '\\') r += '\\'; r += c; }
    return r;
}

static void pair_out(ostream &o, const char *key, const pair<int, int> &p) { o << ",\"" << key << "\":[" << p.first << "," << p.second << "]"; }

int main()
{
    string line;
    while (getline(cin, line)) {
        if (line.empty()) continue;
        vector<string> f;
        { stringstream ss(line); string t; while (getline(ss, t, '\t')) f.push_back(t); }
        if (f.size() < 14) { cerr << "bad case line" << endl; return 2; }
        Sequence s1(f[4], f[1], atoi(f[2].c_str()), atoi(f[3].c_str()) != 0);
        Sequence s2(f[8], f[5], atoi(f[6].c_str()), atoi(f[7].c_str()) != 0);
        const int flag = atoi(f[9].c_str());
        Scorer scr(atoi(f[10].c_str()), atoi(f[11].c_str()), atoi(f[12].c_str()), atoi(f[13].c_str()));
        AGEaligner a(s1, s2);
        ostringstream o;
        o << "{\"id\":\"" << esc(f[0]) << "\"";
        // the reference prints the record itself when the auxiliary aligner of -inv wins (AGEaligner.cpp:181-184):
        // keep stdout for the JSON only
        streambuf *keep = cout.rdbuf();
        ostringstream sink;
        cout.rdbuf(sink.rdbuf());
        const bool ok = a.align(scr, flag);
        if (ok) a.SetAlignResult();
        cout.rdbuf(keep);
        o << ",\"aligned\":" << (ok ? 1 : 0) << ",\"printed_bytes\":" << sink.str().size();
        if (ok) {
            AlignResult r = a.align_result();
            o << ",\"id1\":\"" << esc(r._id1) << "\",\"id2\":\"" << esc(r._id2) << "\",\"score\":" << r._score
              << ",\"strand\":\"" << r._strand << "\",\"is_alternative_align\":" << (r._is_alternative_align ? 1 : 0);
            o << ",\"identity\":[";
            for (size_t i = 0; i < r._identity.size(); ++i) o << (i ? "," : "") << "[" << r._identity[i].first << "," << r._identity[i].second << "]";
            o << "]";
            pair_out(o, "ci_start1", r._ci_start1); pair_out(o, "ci_end1", r._ci_end1);
            pair_out(o, "ci_start2", r._ci_start2); pair_out(o, "ci_end2", r._ci_end2);
            o << ",\"homo_run\":[" << r._homo_run_atbp1 << "," << r._homo_run_atbp2 << "," << r._homo_run_inbp1 << "," << r._homo_run_inbp2
              << "," << r._homo_run_outbp1 << "," << r._homo_run_outbp2 << "]";
            o << ",\"map\":[";
            for (size_t i = 0; i < r._map.size(); ++i) {
                const MapData &m1 = r._map[i].first, &m2 = r._map[i].second;
                o << (i ? "," : "") << "{\"id1\":\"" << esc(m1._id) << "\",\"start1\":" << m1._start << ",\"end1\":" << m1._end
                  << ",\"seq1\":\"" << esc(m1._sequence) << "\",\"id2\":\"" << esc(m2._id) << "\",\"start2\":" << m2._start
                  << ",\"end2\":" << m2._end << ",\"seq2\":\"" << esc(m2._sequence) << "\",\"info\":\"" << esc(r._map_info[i]) << "\"}";
            }
            o << "]";
        }
        o << "}";
        cout << o.str() << "\n";
    }
    return 0;
}
