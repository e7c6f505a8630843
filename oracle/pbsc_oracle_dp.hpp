// pbsc_oracle_dp.hpp — CPU restatement of the DP / multiple-alignment fallback that `stride pbcorrect` runs on
// seed pairs its FM-index walk could not bridge.  TEST INFRASTRUCTURE ONLY (see pbsc_oracle.hpp): the product
// never includes, links or executes this file.
//
// Reference functions restated here (paths relative to /root/reference):
//   LongReadOverlap::retrieveStr             PacBio/LongReadOverlap.cpp:667-756   (read suffixes/prefixes by LF-mapping)
//   LongReadOverlap::retrieveMatches         PacBio/LongReadOverlap.cpp:593-662   (banded alignment + filters)
//   LongReadOverlap::buildMultipleAlignment  PacBio/LongReadOverlap.cpp:17-55
//   Overlapper::extendMatch                  Thirdparty/overlapper.cpp:421-701    (band 200, +1/-1/-8, homopolymer tie-breaks)
//   SequenceOverlap::getPercentIdentity      Thirdparty/overlapper.cpp:71-74
//   MultipleAlignment::_addSequence          Thirdparty/multiple_alignment.cpp:240-393
//   MultipleAlignmentElement::insertGapBeforeColumn  Thirdparty/multiple_alignment.cpp:112-134
//   MultipleAlignment::calculateBaseConsensus        Thirdparty/multiple_alignment.cpp:517-594
// The caller (Corrector::correctByMSAlignment in pbsc_oracle.hpp) follows PacBioSelfCorrectionProcess.cpp:208-245.
//
// Parity is pinned by tests/test_oracle_vs_golden.py: default-option (DP on) correct.fa / discard.fa / summary of the
// reference binary on the golden data sets, and the reference run live when oracle/_ref exists.
#ifndef PBSC_ORACLE_DP_HPP
#define PBSC_ORACLE_DP_HPP

#include <limits>
#include <string>
#include <vector>

namespace pbo {

// ---------------------------------------------------------------------------------
// pairwise overlap, cigar kept expanded (one op per alignment column)
// ---------------------------------------------------------------------------------
struct PairOverlap
{
    int start[2] = {0, 0}, end[2] = {-1, -1};
    int score = -1, editDistance = -1, totalColumns = -1;
    std::string ops;   // 'M', 'I' (base of s2 only), 'D' (base of s1 only), first column first
    // (double)(total_columns - edit_distance) * 100.0f / total_columns
    double percentIdentity() const { return (double)(totalColumns - editDistance) * 100.0f / totalColumns; }
};

// Overlapper::extendMatch.  The score table holds band_width cells per column of s1 and is ZERO-initialised; cells the
// fill loop never visits (column 0, row 0, columns whose band misses the matrix) keep that 0 and are read as such.
inline PairOverlap extendMatch(const std::string& s1, const std::string& s2, int start_1, int start_2, int band_width,
                               const int MATCH_SCORE, const int GAP_PENALTY, const int MISMATCH_PENALTY, bool* ok = nullptr)
{
    PairOverlap out;
    const int nCols = (int)s1.size() + 1, nRows = (int)s2.size() + 1;
    const int half = band_width / 2;
    const int bw = half * 2 + 1;
    const int origin = (start_2 - start_1 + 1) - (half + 1);
    std::vector<int> cells((size_t)nCols * bw, 0);
    // a neighbour outside the band can never be the predecessor: the reference adds a penalty to INT_MIN there, which wraps
    // to a huge positive number and never equals a real score
    const long long OUTSIDE = std::numeric_limits<long long>::min() / 4;
    auto inBand = [&](int i, int j) { const int r = j - (origin + i); return r >= 0 && r < bw; };
    auto at = [&](int i, int j) -> long long { return inBand(i, j) ? (long long)cells[(size_t)i * bw + (j - (origin + i))] : OUTSIDE; };

    for (int i = 1; i < nCols; ++i)
    {
        int j = origin + i, endRow = j + bw;
        if (j < 1) j = 1;
        if (endRow > nRows) endRow = nRows;
        if (endRow <= 0 || j >= nRows || j >= endRow) continue;
        const int first = j, last = endRow - 1;
        for (; j <= last; ++j)
        {
            const long long sub = (s1[i - 1] == s2[j - 1]) ? MATCH_SCORE : MISMATCH_PENALTY;
            const long long diag = (long long)cells[(size_t)(i - 1) * bw + ((j - 1) - (origin + i - 1))] + sub;
            long long v;
            if (j == first)
            {
                // first row of the band: the cell above is never consulted
                v = inBand(i - 1, j) ? std::max(at(i - 1, j) + GAP_PENALTY, diag) : diag;
            }
            else if (j == last)
            {
                // last row: the cell to the left is ignored, even when the band was clipped by the matrix and it exists
                v = std::max(diag, (long long)cells[(size_t)i * bw + ((j - 1) - (origin + i))] + GAP_PENALTY);
            }
            else
            {
                const long long left = (long long)cells[(size_t)(i - 1) * bw + (j - (origin + i - 1))] + GAP_PENALTY;
                const long long up = (long long)cells[(size_t)i * bw + ((j - 1) - (origin + i))] + GAP_PENALTY;
                v = std::max(std::max(diag, left), up);
            }
            cells[(size_t)i * bw + (j - (origin + i))] = (int)v;
        }
    }

    // best cell of the last row (columns 1..) and of the last column (rows 1..); the row wins ties
    long long bestRow = OUTSIDE, bestCol = OUTSIDE;
    size_t bestRowIdx = 0, bestColIdx = 0;
    bool anyRow = false, anyCol = false;
    for (int i = 1; i < nCols; ++i) if (inBand(i, nRows - 1)) { const long long v = at(i, nRows - 1); if (!anyRow || v > bestRow) { bestRow = v; bestRowIdx = i; anyRow = true; } }
    for (int j = 1; j < nRows; ++j) if (inBand(nCols - 1, j)) { const long long v = at(nCols - 1, j); if (!anyCol || v > bestCol) { bestCol = v; bestColIdx = j; anyCol = true; } }
    size_t i, j;
    if (anyCol && (!anyRow || bestCol > bestRow)) { i = nCols - 1; j = bestColIdx; out.score = (int)bestCol; }
    else { i = anyRow ? bestRowIdx : 0; j = nRows - 1; out.score = (int)bestRow; }
    if (ok) *ok = (anyRow || anyCol) && i > 0 && j > 0;   // the reference asserts a non-empty cigar

    out.end[0] = (int)i - 1; out.end[1] = (int)j - 1;
    out.editDistance = 0; out.totalColumns = 0;
    std::string rev;
    while (i > 0 && j > 0)
    {
        const bool isMatch = s1[i - 1] == s2[j - 1];
        const long long diag = at((int)i - 1, (int)j - 1) + (isMatch ? MATCH_SCORE : MISMATCH_PENALTY);
        const long long up = at((int)i, (int)j - 1) + GAP_PENALTY;
        const long long left = at((int)i - 1, (int)j) + GAP_PENALTY;
        const long long curr = at((int)i, (int)j);
        // s2[j] / s1[i] may be one past the end: std::string yields '\0' there
        const char s2next = j < s2.size() ? s2[j] : '\0';
        const char s1next = i < s1.size() ? s1[i] : '\0';
        char op;
        if (s2[j - 1] == s2next) op = (curr == up) ? 'I' : (curr == left) ? 'D' : 'M';          // s2 homopolymer: prefer consuming s2
        else if (s1[i - 1] == s1next) op = (curr == left) ? 'D' : (curr == up) ? 'I' : 'M';     // s1 homopolymer: prefer consuming s1
        else op = (curr == diag) ? 'M' : (curr == left) ? 'D' : 'I';
        if (op == 'M') { if (!isMatch) out.editDistance++; i--; j--; }
        else if (op == 'I') { out.editDistance++; j--; }
        else { out.editDistance++; i--; }
        rev.push_back(op);
        out.totalColumns++;
    }
    out.start[0] = (int)i; out.start[1] = (int)j;
    out.ops.assign(rev.rbegin(), rev.rend());
    return out;
}

// ---------------------------------------------------------------------------------
// LongReadOverlap::retrieveStr: for up to `coverage` suffix-array rows of the seed k-mer on each strand, spell the read
// onwards by LF-mapping until '$' or maxLength symbols
// ---------------------------------------------------------------------------------
inline void retrieveStr(const std::string& query, size_t seedSize, size_t maxLength, const IndexSet& indices, bool isRC, size_t coverage,
                        std::vector<std::string>& ovlStr)
{
    std::string initKmer = isRC ? reverseComplement(query.substr(query.length() - seedSize, seedSize)) : query.substr(0, seedSize);
    BWTInterval fwd = findInterval(indices.pRBWT, reverse(initKmer));
    BWTInterval rvc = findInterval(indices.pBWT, reverseComplement(initKmer));
    for (int64_t root = fwd.lower; fwd.isValid() && root <= fwd.upper && (root - fwd.lower < (int)coverage); root++)
    {
        std::string cur = initKmer;
        int64_t idx = root;
        for (size_t len = initKmer.length(); len < maxLength; len++)
        {
            const char b = indices.pRBWT->getChar(idx);
            if (b == '$') break;
            cur.append(1, b);
            idx = indices.pRBWT->getPC(bwtRank(b)) + indices.pRBWT->occ(bwtRank(b), idx - 1);
            OccCounter::n()++;
        }
        ovlStr.push_back(isRC ? reverseComplement(cur) : cur);
    }
    for (int64_t root = rvc.lower; root <= rvc.upper && rvc.isValid() && (root - rvc.lower < (int)coverage); root++)
    {
        std::string rev;   // symbols in the order LF-mapping yields them (each one precedes the previous in the read)
        int64_t idx = root;
        for (size_t len = initKmer.length(); len < maxLength; len++)
        {
            const char b = indices.pBWT->getChar(idx);
            if (b == '$') break;
            rev.push_back(b);
            idx = indices.pBWT->getPC(bwtRank(b)) + indices.pBWT->occ(bwtRank(b), idx - 1);
            OccCounter::n()++;
        }
        std::string cur(rev.rbegin(), rev.rend());
        cur += reverseComplement(initKmer);
        ovlStr.push_back(isRC ? cur : reverseComplement(cur));
    }
}

struct AlignedRow { std::string seq; PairOverlap ov; };

// LongReadOverlap::retrieveMatches
inline void retrieveMatches(const std::string& query, size_t k, size_t min_overlap, double min_identity, size_t coverage, const IndexSet& indices,
                            bool isRC, std::vector<AlignedRow>& rows, uint64_t* cells = nullptr)
{
    std::vector<std::string> ovlStr;
    const size_t maxLength = query.length() * 1.1 + 20;
    retrieveStr(query, k, maxLength, indices, isRC, coverage, ovlStr);
    for (const std::string& m : ovlStr)
    {
        if ((!isRC && m.substr(0, query.length()) == query) ||
            (isRC && m.length() >= query.length() && m.substr(m.length() - query.length()) == query))
            continue;
        const int bandwidth = 200;
        bool ok = true;
        PairOverlap ov = isRC ? extendMatch(query, m, (int)(query.length() - k), (int)(m.length() - k), bandwidth, 1, -1, -8, &ok)
                              : extendMatch(query, m, 0, 0, bandwidth, 1, -1, -8, &ok);
        if (cells) *cells += (uint64_t)query.length() * 201;
        if (!ok) { fprintf(stderr, "pbsc_oracle: empty alignment (the reference would abort on its assert)\n"); exit(EXIT_FAILURE); }
        const bool passedOverlap = (size_t)ov.totalColumns >= (size_t)min_overlap;
        const bool passedIdentity = ov.percentIdentity() / 100 >= min_identity;
        if (passedOverlap && passedIdentity) rows.push_back(AlignedRow{m, ov});
    }
}

// ---------------------------------------------------------------------------------
// MultipleAlignment restricted to what buildMultipleAlignment + calculateBaseConsensus use: every incoming row is
// placed against row 0 (the query)
// ---------------------------------------------------------------------------------
struct MsaRow
{
    std::string padded;
    size_t leading = 0, trailing = 0;
    size_t numColumns() const { return leading + padded.size() + trailing; }
    char symbol(size_t col) const { return (col < leading || col >= leading + padded.size()) ? '\0' : padded[col - leading]; }
    void insertGapBefore(size_t col)
    {
        if (col <= leading) leading += 1;
        else
        {
            const size_t pos = col - leading;
            if (pos < padded.size()) padded.insert(pos, 1, '-'); else trailing += 1;
        }
    }
};

struct Msa
{
    std::vector<MsaRow> rows;
    void addBase(const std::string& s) { MsaRow r; r.padded = s; rows.push_back(r); }
    void addOverlap(const std::string& seq, const PairOverlap& ov)
    {
        // padded position of base ov.start[0] of row 0
        size_t tpl = 0;
        {
            size_t seen = 0;
            const std::string& p = rows[0].padded;
            for (tpl = 0; tpl < p.size(); ++tpl) if (p[tpl] != '-') { if (seen == (size_t)ov.start[0]) break; seen++; }
        }
        size_t inc = ov.start[1];
        const size_t tplLeading = rows[0].leading;
        const size_t incLeading = tpl + tplLeading;
        std::string out;
        size_t c = 0;
        while (c < ov.ops.size())
        {
            const std::string& p = rows[0].padded;
            const bool inGap = tpl < p.size() && p[tpl] == '-';
            const char op = ov.ops[c];
            if (inGap)
            {
                if (op == 'I') { out.push_back(seq[inc]); inc++; c++; tpl++; }
                else { out.push_back('-'); tpl++; }
            }
            else if (op == 'M') { out.push_back(seq[inc]); inc++; tpl++; c++; }
            else if (op == 'I')
            {
                for (auto& r : rows) r.insertGapBefore(tpl + tplLeading);
                out.push_back(seq[inc]); inc++; c++; tpl++;
            }
            else { out.push_back('-'); c++; tpl++; }   // 'D'
        }
        MsaRow r;
        r.padded = out;
        r.leading = incLeading;
        r.trailing = rows[0].numColumns() - out.size() - incLeading;
        rows.push_back(r);
    }
    // calculateBaseConsensus(min_call_coverage, min_trim_coverage = -1)
    std::string consensus(int min_call_coverage) const
    {
        static const char alphabet[] = "ACGTN-";
        auto index = [](char s) { switch (s) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case '-': return 5; default: return 4; } };
        const MsaRow& base = rows[0];
        const size_t startCol = base.leading, endCol = base.numColumns() - base.trailing - 1;
        std::string cons;
        for (size_t col = startCol; col <= endCol; ++col)
        {
            int counts[6] = {0, 0, 0, 0, 0, 0};
            for (const auto& r : rows) { const char s = r.symbol(col); if (s != '\0') counts[index(s)]++; }
            char maxSym = '\0'; int maxCount = -1;
            for (int a = 0; a < 6; ++a) if (alphabet[a] != 'N' && counts[a] > maxCount) { maxSym = alphabet[a]; maxCount = counts[a]; }
            const char baseSym = base.symbol(col);
            const int baseCount = counts[index(baseSym)];
            const char call = (maxCount >= baseCount && baseCount < min_call_coverage) ? maxSym : baseSym;
            if (call != '-') cons.push_back(call);
        }
        return cons;
    }
};

// LongReadOverlap::buildMultipleAlignment
inline Msa buildMultipleAlignment(const std::string& query, size_t srcK, size_t tarK, size_t min_overlap, double min_identity, size_t coverage,
                                  const IndexSet& indices, uint64_t* cells = nullptr)
{
    Msa msa;
    msa.addBase(query);
    std::vector<AlignedRow> rows;
    retrieveMatches(query, srcK, min_overlap, min_identity, coverage, indices, false, rows, cells);
    retrieveMatches(query, tarK, min_overlap, min_identity, coverage, indices, true, rows, cells);
    for (const auto& r : rows) msa.addOverlap(r.seq, r.ov);
    return msa;
}

}  // namespace pbo
#endif
