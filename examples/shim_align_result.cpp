// AsmvarDetect's use of the aligner (VarUnit.cpp:271-299: align, SetAlignResult, align_result) against
// include/age_shim.hh.  Reads cases from stdin, one per line, tab separated:
//   id  name1  start1  reverse1  seq1  name2  start2  reverse2  seq2  flag  match  mismatch  go  ge
// and prints the AlignResult of each as one JSON object per line (tests/test_shim.py compares them with the
// AlignResult dumps of the reference aligner, tests/golden/align_result_cases.json).
//   g++ -std=c++17 -Iinclude examples/shim_align_result.cpp -Lasmvar_b200 -lage_b200 -Wl,-rpath,$PWD/asmvar_b200 -o shim_align_result
#include "age_shim.hh"

#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace age_b200;
using std::string;

static string q(const string &s)
{
    string r = "\"";
    for (char c : s) { if (c == '"' || c == '\\') r += '\\'; r += c; }
    return r + "\"";
}
static string pr(const std::pair<int, int> &p) { return "[" + std::to_string(p.first) + "," + std::to_string(p.second) + "]"; }

int main()
{
    string line;
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::vector<string> f;
        { std::stringstream ss(line); string t; while (std::getline(ss, t, '\t')) f.push_back(t); }
        if (f.size() < 14) { std::cerr << "bad case line\n"; return 2; }
        Sequence s1(f[4], f[1], atoi(f[2].c_str()), atoi(f[3].c_str()) != 0);
        Sequence s2(f[8], f[5], atoi(f[6].c_str()), atoi(f[7].c_str()) != 0);
        Scorer scr((short)atoi(f[10].c_str()), (short)atoi(f[11].c_str()), (short)atoi(f[12].c_str()), (short)atoi(f[13].c_str()));
        AGEaligner a(s1, s2);
        const bool ok = a.align(scr, atoi(f[9].c_str()));
        std::ostringstream printed;
        if (ok) a.SetAlignResult(printed);
        std::ostringstream o;
        o << "{\"id\":" << q(f[0]) << ",\"aligned\":" << (ok ? 1 : 0) << ",\"printed_bytes\":" << printed.str().size();
        if (ok) {
            const AlignResult r = a.align_result();
            o << ",\"id1\":" << q(r._id1) << ",\"id2\":" << q(r._id2) << ",\"score\":" << r._score << ",\"strand\":\"" << r._strand
              << "\",\"is_alternative_align\":" << (r._is_alternative_align ? 1 : 0) << ",\"identity\":[";
            for (size_t i = 0; i < r._identity.size(); ++i) o << (i ? "," : "") << pr(r._identity[i]);
            o << "],\"ci_start1\":" << pr(r._ci_start1) << ",\"ci_end1\":" << pr(r._ci_end1) << ",\"ci_start2\":" << pr(r._ci_start2)
              << ",\"ci_end2\":" << pr(r._ci_end2) << ",\"homo_run\":[" << r._homo_run_atbp1 << "," << r._homo_run_atbp2 << ","
              << r._homo_run_inbp1 << "," << r._homo_run_inbp2 << "," << r._homo_run_outbp1 << "," << r._homo_run_outbp2 << "],\"map\":[";
            for (size_t i = 0; i < r._map.size(); ++i) {
                const MapData &m1 = r._map[i].first, &m2 = r._map[i].second;
                o << (i ? "," : "") << "{\"id1\":" << q(m1._id) << ",\"start1\":" << m1._start << ",\"end1\":" << m1._end << ",\"seq1\":" << q(m1._sequence)
                  << ",\"id2\":" << q(m2._id) << ",\"start2\":" << m2._start << ",\"end2\":" << m2._end << ",\"seq2\":" << q(m2._sequence)
                  << ",\"info\":" << q(r._map_info[i]) << "}";
            }
            o << "]";
        }
        o << "}";
        std::cout << o.str() << "\n";
    }
    return 0;
}
