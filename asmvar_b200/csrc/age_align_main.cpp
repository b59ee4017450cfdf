// age_align_main.cpp -- the drop-in `age_align` command line on top of the C ABI (include/age_b200.h).
//
// Two dialects, exactly the two the reference ships:
//   batch     (src/AsmvarAGE/src/age_align.cpp:47-297)        age_align [-t -q -s -m -M -g -e -b -a -c -i -v -l -r -d -h] var.txt
//   original  (src/AsmvarAGE/src/Trash/age_align.cpp:36-366)  age_align [-indel|-tdup|-inv|-invl|-invr] [-match=] [-mismatch=] [-go=] [-ge=]
//                                                              [-both] [-revcom1] [-revcom2] [-coor1=s-e] [-coor2=s-e] [-version] f1.fa f2.fa
// The dialect is chosen by the option spelling (`-match=..`, `-coor1=..`, `-indel`, ... are original; getopt
// style short / `--long` options are batch) or forced with AGE_DIALECT=orig|batch.  Where the reference loops over
// candidates and calls AGEaligner::align one at a time, this program collects them and calls age_align_batch();
// stdout keeps the candidate order and the `.age` text byte for byte (the `Alignment time is` line, printed by
// the reference under -DAGE_TIME, reports the batch time divided by the number of candidates).
#include "age_b200.h"

#include <getopt.h>
#include <sys/time.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

namespace {

struct Seq {                                   // the subset of class Sequence the CLI needs (Sequence.hh:20-215)
    string seq, name;
    int start = 1;
    bool reverse = false;
    void revcom() {                            // Sequence.hh:190-198
        if (reverse) start -= (int)seq.length() - 1; else start += (int)seq.length() - 1;
        reverse = !reverse;
        string r(seq.size(), ' ');
        age_revcom(seq.data(), (int32_t)seq.size(), &r[0]);
        seq.swap(r);
    }
};

// Sequence::substr (Sequence.hh:200-215), 1-based, <= 0 = open end.  false = the reference returns NULL here.
bool substr(const string &full, const string &name, int start, int end, Seq &out)
{
    if (start > 0 && end > 0 && start > end) return false;
    const int n = (int)full.length();
    if (start <= 0) start = 1;
    if (end <= 0 || end > n) end = n;
    out.name = name; out.start = start; out.reverse = false;
    out.seq = (n > 0 && start <= n) ? full.substr(start - 1, max(0, end - start + 1)) : string();
    return true;
}

string local_time() { time_t t; time(&t); return asctime(localtime(&t)); }
string upper(string s) { for (auto &c : s) c = (char)toupper((unsigned char)c); return s; }

// gz-aware whole-file reader (the reference uses gzstream, which also reads plain files)
bool slurp(const string &file, string &out)
{
    gzFile f = gzopen(file.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    int n;
    out.clear();
    while ((n = gzread(f, buf, sizeof(buf))) > 0) out.append(buf, n);
    gzclose(f);
    return true;
}

// Fa::Load (Backup/Fa.h:27-50): per line the first whitespace-delimited token; '>' tokens start a record
void load_fa(const string &file, map<string, string> &fa)
{
    string data;
    if (!slurp(file, data)) { cerr << "Cannot open file : " << file << endl; exit(1); }
    string id;
    size_t i = 0;
    const size_t n = data.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)data[i])) ++i;          // operator>> skips leading white space
        if (i >= n) break;
        size_t j = i;
        while (j < n && !isspace((unsigned char)data[j])) ++j;
        if (data[i] == '>') { id.assign(data, i + 1, j - i - 1); fa[id].clear(); }
        else fa[id].append(data, i, j - i);
        while (j < n && data[j] != '\n') ++j;                          // getline drops the rest of the line
        i = j + 1;
    }
}

struct Candidate {                              // one record of the output
    string header;                              // "# info\n" of the batch dialect ("" otherwise)
    Seq s1, s2, s3;                             // s3 = revcom clone of s2 for -both
    bool ok = true;                             // false: sequences could not be extracted
    int job[2] = {-1, -1};
};

struct Options {
    int match = 1, mismatch = -2, go = -2, ge = -1;
    bool both = false, revcom1 = false, revcom2 = false;
    int flag = 0;
};

double now() { timeval t; gettimeofday(&t, nullptr); return t.tv_sec + t.tv_usec / 1e6; }

age_ctx *g_ctx = nullptr;
age_ctx *ctx()
{
    if (g_ctx) return g_ctx;
    vector<int> devs;
    if (const char *e = getenv("AGE_DEVICES")) {   // e.g. AGE_DEVICES=0,1,2,3
        stringstream ss(e); string tok;
        while (getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
    }
    g_ctx = age_create(devs.empty() ? nullptr : devs.data(), (int)devs.size());
    if (!g_ctx) { cerr << "[ERROR] " << age_last_error(nullptr) << endl; exit(1); }
    return g_ctx;
}

// Align a chunk of candidates on the GPU(s) and print their records in order.
void flush_candidates(vector<Candidate> &cands, const Options &o, bool age_time)
{
    if (cands.empty()) return;
    vector<age_job> jobs;
    for (Candidate &c : cands) {
        if (!c.ok) continue;
        auto add = [&](const Seq &a, const Seq &b) {
            age_job j;
            j.seq1 = a.seq.data(); j.len1 = (int32_t)a.seq.size();
            j.seq2 = b.seq.data(); j.len2 = (int32_t)b.seq.size();
            j.flag = o.flag;
            j.match = (int16_t)o.match; j.mismatch = (int16_t)o.mismatch; j.gap_open = (int16_t)o.go; j.gap_extend = (int16_t)o.ge;
            j.partner = -1;
            jobs.push_back(j);
            return (int)jobs.size() - 1;
        };
        c.job[0] = add(c.s1, c.s2);
        if (o.both) {
            c.job[1] = add(c.s1, c.s3);
            jobs[c.job[0]].partner = c.job[1];
            jobs[c.job[1]].partner = c.job[0];
        }
    }
    vector<age_result> res(jobs.size() ? jobs.size() : 1);
    const double t0 = now();
    if (!jobs.empty() && age_align_batch(ctx(), jobs.data(), jobs.size(), res.data()) != 0) {
        cerr << "[ERROR] age_align_batch: " << age_last_error(g_ctx) << endl;
        exit(1);
    }
    const double per = cands.empty() ? 0.0 : (now() - t0) / cands.size();
    string buf;
    for (Candidate &c : cands) {
        fputs(c.header.c_str(), stdout);
        if (c.ok) {
            const age_result *r0 = &res[c.job[0]], *r1 = o.both ? &res[c.job[1]] : nullptr;
            for (const age_result *r : {r0, r1})
                if (r && r->status < 0) {
                    cerr << "[ERROR] alignment of '" << c.s1.name << "' vs '" << c.s2.name << "' is outside the supported range (status "
                         << r->status << "): scores must fit 16 bits, gap penalties must be negative" << endl;
                    exit(1);
                }
            const bool ok0 = r0->status == AGE_OK, ok1 = r1 && r1->status == AGE_OK;
            int pick = -1;
            if (!o.both) pick = ok0 ? 0 : -1;
            else if (ok0 || ok1) pick = (ok0 && (!ok1 || r0->score >= r1->score)) ? 0 : 1;     // age_align.cpp:270-276
            if (pick < 0) cerr << "No alignment made." << endl;
            else {
                const Seq &q = pick ? c.s3 : c.s2;
                const age_result *r = pick ? r1 : r0;
                age_seqview v1{c.s1.seq.data(), (int32_t)c.s1.seq.size(), c.s1.name.c_str(), c.s1.start, c.s1.reverse ? 1 : 0};
                age_seqview v2{q.seq.data(), (int32_t)q.seq.size(), q.name.c_str(), q.start, q.reverse ? 1 : 0};
                const age_job &j = jobs[c.job[pick]];
                buf.resize(8192 + 8 * (c.s1.seq.size() + q.seq.size()));
                size_t n = age_format_alignment(&v1, &v2, &j, r, &buf[0], buf.size());
                if (n >= buf.size()) { buf.resize(n + 1); n = age_format_alignment(&v1, &v2, &j, r, &buf[0], buf.size()); }
                fwrite(buf.data(), 1, n, stdout);
            }
        }
        if (age_time) printf("\nAlignment time is %g s\n\n", per);
    }
    fflush(stdout);
    cands.clear();
}

int mode_count(int flag)
{
    int n = 0;
    for (int f : {AGE_INDEL_FLAG, AGE_TDUPLICATION_FLAG, AGE_INVR_FLAG, AGE_INVL_FLAG, AGE_INVERSION_FLAG}) if (flag & f) ++n;
    return n;
}

const size_t CHUNK = 8192;                      // candidates per age_align_batch call

// ---------------------------------------------------------------------------------------------------------
// batch dialect (age_align.cpp)
// ---------------------------------------------------------------------------------------------------------
void batch_usage(const char *prog, const Options &o, const string &sel)
{
    cerr << "\nAGE -- Alignment with Gap Excision (B200 build, " << age_version() << ")\n\n";
    cerr << "Usage : " << prog << " [Options] [VariantRegionInfile] > Output.age 2> alig.log \n\n";
    cerr << "      Options  : \n";
    cerr << "          -t, --target    [str]  target sequence, fa format. require!\n";
    cerr << "          -q, --query     [str]  Query  sequence, fa format. require!\n";
    cerr << "          -s, --selectId  [str]  Select the target id. [" << sel << "]\n";
    cerr << "          -m, --match     [int]  match score.          [" << o.match << "]\n";
    cerr << "          -M, --mismatch  [int]  mismatch score.       [" << o.mismatch << "]\n";
    cerr << "          -g, --go        [int]  Gap open score.       [" << o.go << "]\n";
    cerr << "          -e, --ge        [int]  Gap Extend score.     [" << o.ge << "]\n";
    cerr << "          -b, --both             Both reverse.         [" << o.both << "]\n";
    cerr << "          -a, --revcom1          revcom1.              [" << o.revcom1 << "]\n";
    cerr << "          -c, --revcom2          revcom2.              [" << o.revcom2 << "]\n";
    cerr << "          -i, --indel            Indel.\n          -v, --inv              Inv.\n          -l, --invl             Invl.\n";
    cerr << "          -d, --tdup             tdup.\n          -r, --invr             Invr.\n";
    cerr << "          -h                    Output this help information.\n\n";
    exit(1);
}

int batch_main(int argc, char **argv)
{
    Options o;
    string varfile, tfile, qfile, sel = "ALL";
    bool indel = false, inv = false, invl = false, invr = false, tdup = false;
    static struct option longopts[] = {
        {"go", 1, nullptr, 'g'}, {"ge", 1, nullptr, 'e'}, {"match", 1, nullptr, 'm'}, {"mismatch", 1, nullptr, 'M'},
        {"indel", 0, nullptr, 'i'}, {"inv", 0, nullptr, 'v'}, {"invl", 0, nullptr, 'l'}, {"invr", 0, nullptr, 'r'},
        {"tdup", 0, nullptr, 'd'}, {"revcom1", 0, nullptr, 'a'}, {"revcom2", 0, nullptr, 'c'}, {"both", 0, nullptr, 'b'},
        {"target", 1, nullptr, 't'}, {"query", 1, nullptr, 'q'}, {"selectId", 1, nullptr, 's'}, {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "g:e:m:M:t:q:s:abcdhilrv", longopts, nullptr)) != -1) {
        switch (c) {
        case 'g': o.go = atoi(optarg); break;
        case 'e': o.ge = atoi(optarg); break;
        case 'm': o.match = atoi(optarg); break;
        case 'M': o.mismatch = atoi(optarg); break;
        case 't': tfile = optarg; break;
        case 'q': qfile = optarg; break;
        case 's': sel = optarg; break;
        case 'a': o.revcom1 = true; break;
        case 'c': o.revcom2 = true; break;
        case 'b': o.both = true; break;
        case 'i': indel = true; break;
        case 'v': inv = true; break;
        case 'l': invl = true; break;
        case 'r': invr = true; break;
        case 'd': tdup = true; break;
        case 'h': batch_usage(argv[0], o, sel); break;
        default: cerr << "\n[ERROR]Unknow option: -" << (char)c << "\n" << endl; exit(1);
        }
    }
    if (argc > 1) varfile = argv[argc - 1];
    if (varfile.empty()) { cerr << "\n[ERROR] Variant in file is empty!\n"; batch_usage(argv[0], o, sel); }
    if (tfile.empty()) { cerr << "\n[ERROR] Target Fa is required!   \n"; batch_usage(argv[0], o, sel); }
    if (qfile.empty()) { cerr << "\n[ERROR] Query  Fa is required!   \n"; batch_usage(argv[0], o, sel); }
    cerr << "\n# [INFO] The parameter of this program:";
    for (int i = 0; i < argc; ++i) cerr << " " << argv[i];
    cerr << "\n" << endl;

    if (indel) o.flag |= AGE_INDEL_FLAG;
    if (inv) o.flag |= AGE_INVERSION_FLAG;
    if (invl) o.flag |= AGE_INVL_FLAG;
    if (invr) o.flag |= AGE_INVR_FLAG;
    if (tdup) o.flag |= AGE_TDUPLICATION_FLAG;
    if (o.flag == 0) o.flag = AGE_INDEL_FLAG;                          // age_align.cpp:205
    if (mode_count(o.flag) != 1) { cerr << "# [Error] In mode specification. More than one mode is specified.\n"; exit(1); }

    map<string, string> tfa, qfa;
    load_fa(tfile, tfa);
    load_fa(qfile, qfa);
    cerr << "# [INFO] Target fa and Query fa are loaded successfully." << local_time();

    // variant table (age_align.cpp:169-190): tId tStart tEnd qId qStart qEnd info, '#' lines skipped
    string data;
    if (!slurp(varfile, data)) { cerr << "# [ERROR] Cannot open file : " << varfile << endl; exit(1); }
    struct Reg { string tid, qid, info; long long ts, te, qs, qe; };
    vector<Reg> regs;
    {
        stringstream in(data);
        string line;
        while (getline(in, line)) {
            stringstream ls(line);
            Reg r;
            if (!(ls >> r.tid)) continue;
            if (r.tid[0] == '#') continue;
            r.ts = r.te = r.qs = r.qe = 0;
            ls >> r.ts >> r.te >> r.qid >> r.qs >> r.qe >> r.info;
            regs.push_back(r);
        }
    }
    cerr << "# [INFO] Variant regions are loaded successfully." << local_time();
    if (regs.empty()) cerr << "[WARNING] The variant list is empty!\n";

    vector<Candidate> cands;
    for (const Reg &r : regs) {
        if (upper(sel) != "ALL" && upper(sel) != upper(r.tid)) continue;
        if (!tfa.count(r.tid)) { flush_candidates(cands, o, true); cerr << "# [ERROR] " << r.tid << " is not in " << tfile << "\n"; exit(1); }
        if (!qfa.count(r.qid)) { flush_candidates(cands, o, true); cerr << "# [ERROR] " << r.qid << " is not in " << qfile << "\n"; exit(1); }
        Candidate c;
        c.header = "# " + r.info + "\n";
        c.ok = substr(tfa[r.tid], r.tid, (int)r.ts, (int)r.te, c.s1) && substr(qfa[r.qid], r.qid, (int)r.qs, (int)r.qe, c.s2);
        if (!c.ok) cerr << "# [ERROR] bad region " << r.tid << ":" << r.ts << "-" << r.te << " / " << r.qid << ":" << r.qs << "-" << r.qe << "\n";
        else {
            if (o.revcom1) c.s1.revcom();
            if (o.revcom2) c.s2.revcom();
            if (o.both) { c.s3 = c.s2; c.s3.revcom(); }
        }
        cands.push_back(std::move(c));
        if (cands.size() >= CHUNK) flush_candidates(cands, o, true);
    }
    flush_candidates(cands, o, true);
    cerr << "\n******************** AGE ALL DONE ********************\n";
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// original dialect (Trash/age_align.cpp)
// ---------------------------------------------------------------------------------------------------------
bool is_nuc(char c) { return strchr("AaTtCcGgUuXxNnRrYyMmKkSsWwBbDdHhVv", c) != nullptr && c != '\0'; }     // Sequence.hh:152-173

// Sequence::parseSequences (Sequence.hh:287-342): name = the whole header line, gaps dropped, bad characters
// discard the record
vector<Seq> parse_sequences(const string &file)
{
    vector<Seq> out;
    ifstream in(file.c_str());
    if (!in) { cerr << "Can't open input file '" << file << "'." << endl; return out; }
    string name, seq, line;
    while (getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '>') {
            if (!name.empty()) { Seq s; s.seq = seq; s.name = name; out.push_back(s); }
            name = line.substr(1); seq.clear();
            continue;
        }
        string checked;
        for (char c : line) {
            if (is_nuc(c)) checked += c;
            else if (!(c == ' ' || c == '-')) {
                cerr << "Unrecognized nucleotide '" << c << "' in file '" << file << "'." << endl;
                seq.clear(); name.clear();
            }
        }
        seq += checked;
    }
    if (!name.empty()) { Seq s; s.seq = seq; s.name = name; out.push_back(s); }
    return out;
}

bool parse_value(const string &str, int &val)                          // Trash/age_align.cpp:36-47
{
    if (str == "-") return false;
    size_t i = 0;
    if (i < str.size() && str[i] == '-') ++i;
    while (i < str.size() && isdigit((unsigned char)str[i])) ++i;
    if (i != str.size()) return false;
    val = (short)strtol(str.c_str(), nullptr, 10);
    return true;
}

bool parse_start_end(const string &str, int &start, int &end)          // Trash/age_align.cpp:49-69
{
    if (str == "-") return false;
    size_t i = 0;
    string s, e;
    while (i < str.size() && isdigit((unsigned char)str[i])) s += str[i++];
    if (i == str.size()) return false;
    if (str[i++] != '-') return false;
    while (i < str.size() && isdigit((unsigned char)str[i])) e += str[i++];
    if (i != str.size()) return false;
    int sv = -1, ev = -1;
    if (!s.empty()) sv = (int)strtol(s.c_str(), nullptr, 10);
    if (!e.empty()) ev = (int)strtol(e.c_str(), nullptr, 10);
    if (sv == 0 || ev == 0) return false;
    if (sv > 0 && ev > 0 && ev < sv) return false;
    start = sv; end = ev;
    return true;
}

int orig_main(int argc, char **argv)
{
    string exe = argv[0];
    exe = exe.substr(exe.rfind('/') + 1);
    const string usage = "Usage:\n\t" + exe + "\n\t\t[-version]\n\t\t[-indel|-tdup|-inv|-invl|-invr]\n\t\t[-match=value:1] [-mismatch=value:-2]\n"
                         "\t\t[-go=value:-2] [-ge=value:-1]\n\t\t[-both] [-revcom1] [-revcom2]\n\t\t[-coor1=start-end] [-coor2=start-end] file1.fa file2.fa";
    const string age_ver = string("AGE ") + age_version() + "\n";
    if (argc == 2 && string(argv[1]) == "-version") { cout << age_ver << endl; return 0; }

    vector<vector<string>> commands;            // one argument vector per command line
    vector<string> lines;
    if (argc > 1) {
        vector<string> a; string line;
        for (int i = 1; i < argc; ++i) { a.push_back(argv[i]); line += " "; line += argv[i]; }
        commands.push_back(a); lines.push_back(line);
    } else {                                    // command lines from standard input (Trash/age_align.cpp:120-158)
        string line;
        while (getline(cin, line)) {
            vector<string> a; stringstream ss(line); string tok;
            while (ss >> tok) a.push_back(tok);
            if (!a.empty()) { commands.push_back(a); lines.push_back(line); }
        }
    }
    if (commands.empty()) { cerr << usage << endl; return 0; }

    int flag = 0;                               // sic: not reset between stdin commands (Trash/age_align.cpp:92)
    string old_file1, old_file2;
    vector<Seq> seqs1, seqs2;
    for (size_t ci = 0; ci < commands.size(); ++ci) {
        const vector<string> &args = commands[ci];
        const string &line = lines[ci];
        Options o;
        string file1, file2;
        int start1 = -1, end1 = -1, start2 = -1, end2 = -1;
        bool error = false, version = false;
        for (const string &a : args) {
            if (a[0] == '-') {
                if (a.compare(0, 7, "-match=") == 0) {
                    if (!parse_value(a.substr(7), o.match)) cerr << line << "\nError in match specification '" << a << "'.\nUsing default value " << o.match << "." << endl;
                } else if (a.compare(0, 10, "-mismatch=") == 0) {
                    if (!parse_value(a.substr(10), o.mismatch)) cerr << line << "\nError in mismatch specification '" << a << "'.\nUsing default value " << o.mismatch << "." << endl;
                } else if (a.compare(0, 4, "-go=") == 0) {
                    if (!parse_value(a.substr(4), o.go)) cerr << line << "\nError in gap open specification '" << a << "'.\nUsing default value " << o.go << "." << endl;
                } else if (a.compare(0, 4, "-ge=") == 0) {
                    if (!parse_value(a.substr(4), o.ge)) cerr << line << "\nError in gap extend specification '" << a << "'.\nUsing default value " << o.ge << "." << endl;
                } else if (a.compare(0, 7, "-coor1=") == 0) {
                    if (!parse_start_end(a.substr(7), start1, end1)) { cerr << line << "\nError in coordinate specification '" << a << "'.\n" << usage << endl; error = true; break; }
                } else if (a.compare(0, 7, "-coor2=") == 0) {
                    if (!parse_start_end(a.substr(7), start2, end2)) { cerr << line << "\nError in coordinate specification '" << a << "'.\n" << usage << endl; error = true; break; }
                } else if (a == "-version") version = true;
                else if (a == "-indel") flag |= AGE_INDEL_FLAG;
                else if (a == "-inv") flag |= AGE_INVERSION_FLAG;
                else if (a == "-invl") flag |= AGE_INVL_FLAG;
                else if (a == "-invr") flag |= AGE_INVR_FLAG;
                else if (a == "-tdup") flag |= AGE_TDUPLICATION_FLAG;
                else if (a == "-revcom1") o.revcom1 = true;
                else if (a == "-revcom2") o.revcom2 = true;
                else if (a == "-both") o.both = true;
                else { cerr << line << "\nUnknown option '" << a << "'.\n" << usage << endl; error = true; break; }
                continue;
            }
            if (file1.empty()) file1 = a;
            else if (file2.empty()) file2 = a;
            else { cerr << line << "\nToo many input files '" << a << "'.\n" << usage << endl; error = true; break; }
        }
        if (version) { cout << age_ver << endl; return 0; }
        if (flag == 0) flag = AGE_INDEL_FLAG;
        if (mode_count(flag) != 1) { cerr << "Error in mode specification. More than one mode is specified." << endl; error = true; }
        if (!file1.empty() && file2.empty()) cerr << "No second file given." << endl;
        if (error || file1.empty() || file2.empty()) continue;
        o.flag = flag;
        if (file1 != old_file1) { seqs1 = parse_sequences(file1); old_file1 = file1; }
        if (file2 != old_file2) { seqs2 = parse_sequences(file2); old_file2 = file2; }
        vector<Candidate> cands;
        for (const Seq &a : seqs1)
            for (const Seq &b : seqs2) {
                Candidate c;
                c.ok = substr(a.seq, a.name, start1, end1, c.s1) && substr(b.seq, b.name, start2, end2, c.s2);
                if (c.ok) {
                    if (o.revcom1) c.s1.revcom();
                    if (o.revcom2) c.s2.revcom();
                    if (o.both) { c.s3 = c.s2; c.s3.revcom(); }
                }
                cands.push_back(std::move(c));
                if (cands.size() >= CHUNK) flush_candidates(cands, o, true);
            }
        flush_candidates(cands, o, true);
    }
    return 1;                                   // sic: the reference's original CLI exits with 1 (Trash/age_align.cpp:365)
}

bool looks_original(int argc, char **argv)
{
    if (const char *e = getenv("AGE_DIALECT")) return string(e) == "orig";
    // the batch dialect cannot run without its target / query options; everything else is the original dialect
    // (which also reads command lines from standard input when it has no arguments)
    for (int i = 1; i < argc; ++i) {
        const string a = argv[i];
        if (a == "-t" || a == "-q" || a == "-h" || a == "--target" || a == "--query" || a.compare(0, 9, "--target=") == 0 || a.compare(0, 8, "--query=") == 0)
            return false;
    }
    return true;
}

} // namespace

int main(int argc, char **argv)
{
    const int rc = looks_original(argc, argv) ? orig_main(argc, argv) : batch_main(argc, argv);
    if (g_ctx) age_destroy(g_ctx);
    return rc;
}
