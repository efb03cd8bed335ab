// Enumeration of the dirty-map words the snapshot ring pass has to read (shared by the imprint kernel and a host
// test hook, so that the index arithmetic the device runs is the arithmetic the CPU test checks).
//
// updateSnapshot(canvas, centre) (FootprintBrush.hxx:278-319) refreshes the ring "allowed box minus open interior of
// the footprint box". The dirty map is scanned in 32-bit words (4 pixels). Rows above and below the footprint box need
// every word of the allowed width; rows crossing the box only need the words to the left and to the right of the
// words lying completely inside the open interior. Enumerating exactly those words (instead of the whole allowed
// rectangle with a skip test) keeps large footprints at one or two word loads per thread.
#pragma once

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

struct RingGeom {
  int tlx, tly, brx, bry;  // footprint box corners (exclusive interior bounds)
  int ax0, ay0, ax1, ay1;  // allowed box, clipped to canvas and stored rows
};

struct RingWords {
  int total;  // number of words to read; word i = at(i)
  int w0, nw;           // first word and word count of a full row
  int n_top_words;      // words of the rows above the box (full width)
  int n_full_words;     // + words of the rows below the box
  int row_top, row_mid, row_bot;  // first row of each group
  int nl, right0, nm;   // box rows: nl words from w0, then words from right0; nm = words per box row
  float inv_nw, inv_nm;

  PB_HD explicit RingWords(const RingGeom& g) {
    total = 0;
    w0 = nw = n_top_words = n_full_words = row_top = row_mid = row_bot = nl = right0 = nm = 0;
    inv_nw = inv_nm = 0.0f;
    if (g.ax1 < g.ax0 || g.ay1 < g.ay0) return;
    w0              = g.ax0 >> 2;
    nw              = (g.ax1 >> 2) - w0 + 1;
    const int nrows = g.ay1 - g.ay0 + 1;
    const int iw0 = (g.tlx >> 2) + 1, iw1 = (g.brx - 4) >> 2;  // words completely inside the open interior
    const int bot_first = (g.bry > g.tly + 1 ? g.bry : g.tly + 1);
    int n_top = g.tly - g.ay0 + 1;
    n_top     = n_top < 0 ? 0 : (n_top > nrows ? nrows : n_top);
    int n_bot = g.ay1 - bot_first + 1;
    n_bot     = n_bot < 0 ? 0 : (n_bot > nrows - n_top ? nrows - n_top : n_bot);
    const int n_mid = nrows - n_top - n_bot;
    nl              = iw0 - w0;
    nl              = nl < 0 ? 0 : (nl > nw ? nw : nl);
    right0          = iw1 + 1 > w0 + nl ? iw1 + 1 : w0 + nl;
    int nr          = w0 + nw - right0;
    nr              = nr < 0 ? 0 : nr;
    nm              = nl + nr;
    row_top         = g.ay0;
    row_mid         = g.ay0 + n_top;
    row_bot         = g.ay1 - n_bot + 1;
    n_top_words     = n_top * nw;
    n_full_words    = n_top_words + n_bot * nw;
    total           = n_full_words + n_mid * nm;
    inv_nw          = 1.0f / static_cast<float>(nw);
    inv_nm          = nm > 0 ? 1.0f / static_cast<float>(nm) : 0.0f;
  }

  // word i -> (canvas row, word index within the row); i in [0, total)
  PB_HD void at(int i, int& row, int& wi) const {
    const bool mid  = i >= n_full_words;
    const bool bot  = !mid && i >= n_top_words;
    const int j     = i - (mid ? n_full_words : (bot ? n_top_words : 0));
    const int d     = mid ? nm : nw;
    const float inv = mid ? inv_nm : inv_nw;
    int r = static_cast<int>((static_cast<float>(j) + 0.5f) * inv);  // j / d without integer division, then corrected
    if (r * d > j) --r;
    if ((r + 1) * d <= j) ++r;
    const int c = j - r * d;
    row         = (mid ? row_mid : (bot ? row_bot : row_top)) + r;
    wi          = mid ? (c < nl ? w0 + c : right0 + (c - nl)) : w0 + c;
  }
};

}  // namespace pb
