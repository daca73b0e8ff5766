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
};

constexpr int SPMM_MAX_STAGES = 4;

}  // namespace hmcb
