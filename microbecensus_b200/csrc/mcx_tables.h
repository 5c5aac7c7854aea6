// mcx_tables.h -- scoring tables and constants shared by the host and device code of libmcx.
//
// Every constant is pinned to the prebuilt RAPsearch2 v2.15 binary the reference shells out to
// (/root/reference/microbe_census/bin/rapsearch_Linux_2.15, call site microbe_census.py:375);
// addresses are virtual addresses in that binary.  DESIGN.md has the full derivation.
#pragma once
#include <stdint.h>

namespace mcx {

constexpr int AA_STOP = 20;        // '.', SEG-masked 'x', database 'X': scores -5 against everything
constexpr int MAX_FRAME = 168;     // aa per frame at 500 bp
constexpr int MAX_LINES = 500;     // RAPsearch2 -v default: lines printed per query
constexpr int GAP_SLACK = 63;      // subject columns beyond the query length in a gapped extension
constexpr int GAP_OPEN = 11;       // CHashSearch +0x40358
constexpr int GAP_EXT = 1;         // CHashSearch +0x4035c
constexpr int SEED_MIN_SCORE = 11; // CHashSearch +0x403a0
constexpr int SEED_MIN_IDENT = 4;  // CHashSearch +0x403a8
constexpr int UNGAP_FLOOR = -20;   // AlignFwd/AlignBwd 0x406d6c

// CHashSearch::Search 0x418cee-0x418dca: bit scores through raw = (bits*ln2 + ln K)/lambda,
// BLOSUM62 Karlin-Altschul sets from BlastStat::SetPar (ungapped 0.318/0.134, gapped 11/1 0.267/0.041)
constexpr double LN2 = 0.6931471805599453;
constexpr double UNGAP_XDROP = (7.0 * LN2 + -2.0099154790312257) / 0.318;   //  8.94
constexpr double GAP_TRIGGER = (25.0 * LN2 + -2.0099154790312257) / 0.318;  // 48.17
constexpr double GAP_XDROP = (15.0 * LN2 + -3.1941832122778293) / 0.267;    // 26.98

// residue order ARNDCQEGHILKMFPSTWYV; symbol `blosum62` (.data 0x6749e0)
static const int8_t BLOSUM62[20][20] = {
 { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0},
 {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3},
 {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3},
 {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3},
 { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1},
 {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2},
 {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2},
 { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3},
 {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3},
 {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3},
 {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1},
 {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2},
 {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1},
 {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1},
 {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2},
 { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2},
 { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0},
 {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3},
 {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1},
 { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4}};

// murphy10 groups A | KR | EDNQ | C | G | H | ILVM | FYW | P | ST (symbol `murphy10`, .data 0x67f2c0)
static const uint8_t MURPHY10[21] = {0,1,2,2,3,2,2,4,5,6,6,1,6,7,8,9,9,7,7,6,10};
constexpr uint8_t MURPHY10_CE[21] = {0,1,2,2,3,2,2,4,5,6,6,1,6,7,8,9,9,7,7,6,10};

// codon table, index 16*b0+4*b1+b2 with T=0 C=1 A=2 G=3 (symbol `aa`, .data 0x688c00)
static const char CODON_AA[65] = "FFLLSSSSYY..CC.WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
static const char AA_ORDER[21] = "ARNDCQEGHILKMFPSTWYV";

// seed words: 9 contiguous murphy10 letters, or a 10-letter window whose letter at offset 3..6 is
// replaced (CHashSearch::Searching 0x415050: exact seeds are 9 long because the median bucket size
// of the database is 0; substitution multipliers {10,1,100} and the first-extra-letter loop 0x416365)
constexpr int N_PAT = 5;
static const int PAT_LEN[N_PAT] = {9, 10, 10, 10, 10};
static const int PAT_WILD[N_PAT] = {-1, 3, 4, 5, 6};

// SEG as ported into RAPsearch2 (class Seg): window 12, locut 2.2, hicut 2.5, maxtrim 100,
// downset 0 / upset 1 (Seg::initialize 0x439650 never recomputes them)
constexpr int SEG_WINDOW = 12;
constexpr double SEG_LOCUT = 2.2;
constexpr double SEG_HICUT = 2.5;
constexpr int SEG_MAXTRIM = 100;

}  // namespace mcx
