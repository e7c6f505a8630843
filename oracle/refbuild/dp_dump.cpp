// TEST INFRASTRUCTURE ONLY (oracle/): a tiny driver linked against the reference's own objects, exposing the two
// third-party pieces of the DP fallback that the reference never prints:
//   Overlapper::extendMatch            (Thirdparty/overlapper.cpp:421-701)
//   MultipleAlignment::addOverlap + calculateBaseConsensus   (Thirdparty/multiple_alignment.cpp:215-393, 517-594)
//
//   dp_dump FILE
// FILE holds records:
//   A s1 s2 start1 start2          -> "A score s1start s1end s2start s2end edit columns cigar"
//   M query minCall n              followed by n lines "seq start1 start2": every row is aligned with extendMatch
//                                  (band 200, +1/-1/-8) and added with addOverlap  -> "M rows consensus"
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include "multiple_alignment.h"
#include "overlapper.h"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: dp_dump FILE\n"); return 2; }
    std::ifstream in(argv[1]);
    std::string tag;
    while (in >> tag)
    {
        if (tag == "A")
        {
            std::string s1, s2; int a, b;
            in >> s1 >> s2 >> a >> b;
            SequenceOverlap o = Overlapper::extendMatch(s1, s2, a, b, 200, 1, -1, -8);
            std::cout << "A " << o.score << " " << o.match[0].start << " " << o.match[0].end << " " << o.match[1].start << " " << o.match[1].end << " "
                      << o.edit_distance << " " << o.total_columns << " " << o.cigar << "\n";
        }
        else if (tag == "M")
        {
            std::string q; int minCall, n;
            in >> q >> minCall >> n;
            MultipleAlignment ma;
            ma.addBaseSequence("query", q, "");
            for (int i = 0; i < n; i++)
            {
                std::string s; int a, b;
                in >> s >> a >> b;
                SequenceOverlap o = Overlapper::extendMatch(q, s, a, b, 200, 1, -1, -8);
                ma.addOverlap("row", s, "", o);
            }
            std::cout << "M " << ma.getNumRows() << " " << ma.calculateBaseConsensus(minCall, -1) << "\n";
        }
    }
    return 0;
}
