// top_entries.cuh — index helpers of the 13x13 (host,target) top block (AccumulatorApprox, MatrixAccumulators.h:744-1170):
// 91 upper-triangle entries = 55 of the 10x10 part (row-major), 30 TopRight (10x3), 6 BotRight; entry 91 of a table row is
// the residual count.  13 = 4 calibration + 6 pose + 2 affine + 1 residual.
#pragma once

// entry e of the 91 (10x10 upper triangle row-major, then 10x3 TopRight, then 6 BotRight)
struct EntryDesc { int p, q, kind; };  // kind 0: 10x10 (r=p,c=q)  1: TopRight (i=p,j=q)  2: BotRight (k=p)  3: none
__device__ __forceinline__ EntryDesc entry_desc(int e) {
  EntryDesc d;
  if (e < 55) {
    int r = 0, base = 0;
    while (e >= base + (10 - r)) { base += 10 - r; r++; }
    d.p = r; d.q = r + (e - base); d.kind = 0;
  } else if (e < 85) { d.p = (e - 55) / 3; d.q = (e - 55) % 3; d.kind = 1; }
  else if (e < 91) { d.p = e - 85; d.q = 0; d.kind = 2; }
  else { d.p = d.q = 0; d.kind = 3; }
  return d;
}


// (row, column) of entry e < 91 inside the symmetric 13x13 block, row <= column
__device__ __forceinline__ void entry_rc(int e, int &r, int &c) {
  const EntryDesc d = entry_desc(e);
  if (d.kind == 0) { r = d.p; c = d.q; }
  else if (d.kind == 1) { r = d.p; c = 10 + d.q; }
  else { r = d.p < 3 ? 10 : d.p < 5 ? 11 : 12; c = d.p < 3 ? 10 + d.p : d.p < 5 ? 8 + d.p : 12; }   // BotRight: (10,10) (10,11) (10,12) (11,11) (11,12) (12,12)
}
