// spmm_types.cuh -- tables of the shared-memory staged CSR SpMM (kernel: spmm_strip.cuh; built by
// upload_csr_strips in hmcb.cu).
#pragma once
#include <cuda.h>   // CUtensorMap

namespace hmcb {

// One nonzero of a (row chunk, column strip) group.  The group is stored warp by warp: a header
// of `hdr` 16-byte slots holding the first slot of every consumer warp, then for each warp the
// nonzeros of its RW rows in row order followed by a sentinel (row = RW).
struct SpmmEntry {
  double val;
  int off;   // local column x S x 8: byte offset of the B row inside the staged strip
  int row;   // row inside the warp's RW rows; RW marks the end of the warp's stream
};
static_assert(sizeof(SpmmEntry) == 16, "SpmmEntry must be 16 bytes");

// Compact form, used when every value of the matrix is exactly representable in fp32 (the
// reference rounds G to numpy.single by default, LinearMatrix.py:323): 8 bytes per nonzero.
struct SpmmEntry32 {
  float val;
  unsigned meta;   // bits 0-23: byte offset of the B row inside the staged strip, bits 24-31: row
};
static_assert(sizeof(SpmmEntry32) == 8, "SpmmEntry32 must be 8 bytes");

struct SpmmStrip {
  int col0, ncols;      // B rows col0 + j * cstride, j < ncols
  int ent_off, ent_cnt; // 16-byte units: offset and size of this (chunk, strip) group in the packed array
};

struct StripDev {
  const void* ent;          // SpmmEntry[] or SpmmEntry32[] (compact)
  int compact;              // 1: 8-byte nonzeros
  const SpmmStrip* strips;
  const int* strip_ptr;     // [chunks + 1]
  int rows, chunks;
  int cstride;              // column stride of a strip (= strips per chunk): strips interleave the columns
  int warps, rw, cpl;       // thread mapping the tables were built for
  int kb, emax, stages;     // strip limits (columns, 16-byte slots) and pipeline depth
  int b_bytes, stage_bytes;
  int kb_box;               // rows of the TMA box = widest strip = ceil(cols / cstride) <= kb
  // row-blocked tensor-core form (csr_spmm_block_kernel): rows are regrouped into groups of
  // SPMM_BLOCK_R = 8 rows with similar column sets (one DMMA M-tile), `gw` groups per consumer warp;
  // `ent` then holds k-tiles of 4 columns with their 8 x 4 A fragments
  int blocked;              // 1: the tables below are in the row-blocked format
  const int* perm;          // [chunks x warps x gw x SPMM_BLOCK_R] original row of every block row, -1 = padding
  int gw;                   // row groups per consumer warp
  int nb;                   // 16-chain TMA boxes per slab (slab = 16 nb chains = 2 nb DMMA N-tiles)
  int box_rows, row_boxes;  // a strip is staged as row_boxes x nb boxes of box_rows rows x 128 bytes
};

constexpr int SPMM_MAX_STAGES = 4;

// Row-blocked tables.  A (chunk, strip) group is
//   header  : (warps x gw + 1) u32 -- per (warp, group) (first word of its record stream << 10 | number
//             of k-tiles), then the group's total number of k-tiles; padded to 16 bytes
//   records : per (warp, group) one k-tile record after the other:
//               4 x {u32 w, u32 n}   per column c of the k-tile: w = (byte offset of its B row inside the
//                                    staged strip) / 16 | (mask of the rows with a nonzero in this column) << 13
//                                    | (number of values of the columns before c) << 21; n = number of
//                                    values of the whole k-tile (the same in all four)
//               values               the nonzeros column by column, rows ascending; fp32 when every value of
//                                    the matrix is exact in fp32 (`compact`), else fp64; padded to 8 bytes
//             (the 8 x 4 fragments are 30-45 % dense: storing only the nonzeros halves the bytes streamed)
// B row offsets carry the 128-byte TMA swizzle of their row: row_box * nb * box_rows * 128 + r * 128 +
// (r & 7) * 16 (r = row inside the box).  The chains of a 16-chain box are split over two N-tiles as
// even / odd chains, so lane (k, n) of the B fragment needs chains 2n and 2n + 1 of row k: one 16-byte
// load at offset ^ (n * 16).  The 4 columns of a k-tile come from 4 different classes (r >> 1) & 3
// whenever possible: the quarter warps then read 4 x 32 bytes in 4 different bank groups.
constexpr int SPMM_BLOCK_R = 8;

}  // namespace hmcb
