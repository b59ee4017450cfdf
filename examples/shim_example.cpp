// The candidate loop of the reference's age_align.cpp:244-283, written against include/age_shim.hh:
// two FASTA-less sequences from the command line, -indel -both, scoring 1/-1/-10/-1.
//   g++ -std=c++17 -Iinclude examples/shim_example.cpp -Lasmvar_b200 -lage_b200 -Wl,-rpath,$PWD/asmvar_b200 -o shim_example
#include "age_shim.hh"
#include <iostream>
using namespace age_b200;

int main(int argc, char **argv)
{
    if (argc < 3) { std::cerr << "usage: shim_example SEQ1 SEQ2\n"; return 2; }
    Sequence seq1(argv[1], "first", 1, false), seq2(argv[2], "second", 1, false);
    Scorer scr(1, -1, -10, -1);
    const int flag = AGEaligner::INDEL_FLAG;
    Sequence *seq3 = seq2.clone();
    seq3->revcom();
    AGEaligner aligner1(seq1, seq2), aligner2(seq1, *seq3);
    const bool res1 = aligner1.align(scr, flag), res2 = aligner2.align(scr, flag);
    if (!res1 && !res2) std::cerr << "No alignment made." << std::endl;
    else if (aligner1.score() >= aligner2.score()) aligner1.printAlignment();
    else aligner2.printAlignment();
    delete seq3;
    return 0;
}
