// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// PARITY PINNED for this file: tests/test_ref_pin.py runs every class below against the reference class itself, compiled
// unmodified from /root/reference/src/OptimizationBackend/{MatrixAccumulators,ScaleAccumulator}.h (oracle/ref_build.sh), on
// random streams that cross both shiftUp tiers — bit-identical results.
//
// Restatement of src/OptimizationBackend/MatrixAccumulators.h and ScaleAccumulator.h: the
// numerically tiered (1 / 1k / 1M) float accumulators.  4-lane SSE members are restated as
// float[4] lanes (same lane-wise arithmetic; the parity build uses -ffp-contract=off).
#pragma once
#include <cstring>
#include <cstddef>

namespace orc {

// MatrixAccumulators.h:744-1170.  H = [x y] * [a b; b c] * [x y]^T for the 10x10 part.
struct AccumulatorApprox {
  float H[13][13];
  size_t num;
  float Data[60], Data1k[60], Data1m[60];
  float TopRight_Data[32], TopRight_Data1k[32], TopRight_Data1m[32];
  float BotRight_Data[8], BotRight_Data1k[8], BotRight_Data1m[8];
  float numIn1, numIn1k, numIn1m;

  void initialize() {  // :752-765
    memset(Data, 0, sizeof(Data)); memset(Data1k, 0, sizeof(Data1k)); memset(Data1m, 0, sizeof(Data1m));
    memset(TopRight_Data, 0, sizeof(TopRight_Data)); memset(TopRight_Data1k, 0, sizeof(TopRight_Data1k));
    memset(TopRight_Data1m, 0, sizeof(TopRight_Data1m));
    memset(BotRight_Data, 0, sizeof(BotRight_Data)); memset(BotRight_Data1k, 0, sizeof(BotRight_Data1k));
    memset(BotRight_Data1m, 0, sizeof(BotRight_Data1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void finish() {  // :766-794
    memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int c = r; c < 10; c++) { H[r][c] = H[c][r] = Data1m[idx]; idx++; }
    idx = 0;
    for (int r = 0; r < 10; r++)
      for (int c = 0; c < 3; c++) { H[r][c + 10] = H[c + 10][r] = TopRight_Data1m[idx]; idx++; }
    H[10][10] = BotRight_Data1m[0];
    H[10][11] = H[11][10] = BotRight_Data1m[1];
    H[10][12] = H[12][10] = BotRight_Data1m[2];
    H[11][11] = BotRight_Data1m[3];
    H[11][12] = H[12][11] = BotRight_Data1m[4];
    H[12][12] = BotRight_Data1m[5];
    num = (size_t)(numIn1 + numIn1k + numIn1m);
  }
  // :928-1055 — x = (x4,x6), y = (y4,y6); upper triangle of the 10x10, row by row
  void update(const float *x4, const float *x6, const float *y4, const float *y6, float a, float b, float c) {
    float x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int cc = r; cc < 10; cc++) {
        // Data[idx] += a*x[cc]*x[r] + c*y[cc]*y[r] + b*(x[cc]*y[r] + y[cc]*x[r])
        Data[idx] += a * x[cc] * x[r] + c * y[cc] * y[r] + b * (x[cc] * y[r] + y[cc] * x[r]);
        idx++;
      }
    num++; numIn1++;
    shiftUp(false);
  }
  // :1057-1101
  void updateTopRight(const float *x4, const float *x6, const float *y4, const float *y6, float TR00, float TR10,
                      float TR01, float TR11, float TR02, float TR12) {
    for (int i = 0; i < 10; i++) {
      const float xv = i < 4 ? x4[i] : x6[i - 4], yv = i < 4 ? y4[i] : y6[i - 4];
      TopRight_Data[3 * i + 0] += xv * TR00 + yv * TR10;
      TopRight_Data[3 * i + 1] += xv * TR01 + yv * TR11;
      TopRight_Data[3 * i + 2] += xv * TR02 + yv * TR12;
    }
  }
  // :1103-1112
  void updateBotRight(float a00, float a01, float a02, float a11, float a12, float a22) {
    BotRight_Data[0] += a00; BotRight_Data[1] += a01; BotRight_Data[2] += a02;
    BotRight_Data[3] += a11; BotRight_Data[4] += a12; BotRight_Data[5] += a22;
  }
  // :1129-1169
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1k[i] = Data[i] + Data1k[i];
      for (int i = 0; i < 32; i++) TopRight_Data1k[i] = TopRight_Data[i] + TopRight_Data1k[i];
      for (int i = 0; i < 8; i++) BotRight_Data1k[i] = BotRight_Data[i] + BotRight_Data1k[i];
      numIn1k += numIn1; numIn1 = 0;
      memset(Data, 0, sizeof(Data)); memset(TopRight_Data, 0, sizeof(TopRight_Data)); memset(BotRight_Data, 0, sizeof(BotRight_Data));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1m[i] = Data1k[i] + Data1m[i];
      for (int i = 0; i < 32; i++) TopRight_Data1m[i] = TopRight_Data1k[i] + TopRight_Data1m[i];
      for (int i = 0; i < 8; i++) BotRight_Data1m[i] = BotRight_Data1k[i] + BotRight_Data1m[i];
      numIn1m += numIn1k; numIn1k = 0;
      memset(Data1k, 0, sizeof(Data1k)); memset(TopRight_Data1k, 0, sizeof(TopRight_Data1k)); memset(BotRight_Data1k, 0, sizeof(BotRight_Data1k));
    }
  }
};

// MatrixAccumulators.h:33-78 : A += w * L * R^T   (I x J, row-major here)
template <int I, int J> struct AccumulatorXX {
  float A[I * J], A1k[I * J], A1m[I * J];
  size_t num;
  float numIn1, numIn1k, numIn1m;
  void initialize() { memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m)); num = 0; numIn1 = numIn1k = numIn1m = 0; }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
  void update(const float *L, const float *R, float w) {
    for (int i = 0; i < I; i++)
      for (int j = 0; j < J; j++) A[i * J + j] += (w * L[i]) * R[j];  // Eigen: (w*L) * R^T evaluated coefficient-wise
    numIn1++;
    shiftUp(false);
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < I * J; i++) { A1k[i] += A[i]; A[i] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int i = 0; i < I * J; i++) { A1m[i] += A1k[i]; A1k[i] = 0; } numIn1m += numIn1k; numIn1k = 0; }
  }
};

// MatrixAccumulators.h:152-202 : A += w * L
template <int I> struct AccumulatorX {
  float A[I], A1k[I], A1m[I];
  size_t num;
  float numIn1, numIn1k, numIn1m;
  void initialize() { memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m)); num = 0; numIn1 = numIn1k = numIn1m = 0; }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
  void update(const float *L, float w) { for (int i = 0; i < I; i++) A[i] += w * L[i]; numIn1++; shiftUp(false); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < I; i++) { A1k[i] += A[i]; A[i] = 0; } numIn1k += numIn1; numIn1 = 0; }
    if (numIn1k > 1000 || force) { for (int i = 0; i < I; i++) { A1m[i] += A1k[i]; A1k[i] = 0; } numIn1m += numIn1k; numIn1k = 0; }
  }
};

// MatrixAccumulators.h:1172-1687 : 9x9 upper triangle, 4 SSE lanes per entry.
// MatrixAccumulators.h:80-149 (only the scalar update is used on this path)
struct Accumulator11 {
  float A;
  size_t num;
  float SSEData[4], SSEData1k[4], SSEData1m[4];
  float numIn1, numIn1k, numIn1m;
  void initialize() { A = 0; memset(SSEData, 0, sizeof(SSEData)); memset(SSEData1k, 0, sizeof(SSEData1k)); memset(SSEData1m, 0, sizeof(SSEData1m)); num = 0; numIn1 = numIn1k = numIn1m = 0; }
  void finish() { shiftUp(true); A = SSEData1m[0] + SSEData1m[1] + SSEData1m[2] + SSEData1m[3]; }
  void updateSingle(float val) { SSEData[0] += val; num++; numIn1++; shiftUp(false); }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < 4; i++) SSEData1k[i] = SSEData[i] + SSEData1k[i]; numIn1k += numIn1; numIn1 = 0; memset(SSEData, 0, sizeof(SSEData)); }
    if (numIn1k > 1000 || force) { for (int i = 0; i < 4; i++) SSEData1m[i] = SSEData1k[i] + SSEData1m[i]; numIn1m += numIn1k; numIn1k = 0; memset(SSEData1k, 0, sizeof(SSEData1k)); }
  }
};

struct Accumulator9 {
  float H[9][9];
  size_t num;
  float SSEData[4 * 45], SSEData1k[4 * 45], SSEData1m[4 * 45];
  float numIn1, numIn1k, numIn1m;
  void initialize() { memset(H, 0, sizeof(H)); memset(SSEData, 0, sizeof(SSEData)); memset(SSEData1k, 0, sizeof(SSEData1k)); memset(SSEData1m, 0, sizeof(SSEData1m)); num = 0; numIn1 = numIn1k = numIn1m = 0; }
  void finish() {  // :1189-1205
    memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        float d = SSEData1m[idx + 0] + SSEData1m[idx + 1] + SSEData1m[idx + 2] + SSEData1m[idx + 3];
        H[r][c] = H[c][r] = d;
        idx += 4;
      }
  }
  // :1206-1312 ; J[k][lane]
  void updateSSE(const float J[9][4]) {
    float *pt = SSEData;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        for (int l = 0; l < 4; l++) pt[l] = pt[l] + J[r][l] * J[c][l];
        pt += 4;
      }
    num += 4; numIn1++;
    shiftUp(false);
  }
  // :1543-1660 (lane `off` = 0): diagonal (Jr * Jr) * w, then Jr *= w and the rest of the row Jc * Jr
  void updateSingleWeighted(const float Jin[9], float w) {
    float J[9];
    for (int i = 0; i < 9; i++) J[i] = Jin[i];
    float *pt = SSEData;
    for (int r = 0; r < 9; r++) {
      *pt += J[r] * J[r] * w;
      pt += 4;
      J[r] *= w;
      for (int c = r + 1; c < 9; c++) { *pt += J[c] * J[r]; pt += 4; }
    }
    num++; numIn1++;
    shiftUp(false);
  }
  // :1314-1432 ; J[k][lane], w[lane]
  void updateSSE_eighted(const float J[9][4], const float w[4]) {
    float *pt = SSEData;
    for (int r = 0; r < 9; r++) {
      float Jw[4];
      for (int l = 0; l < 4; l++) Jw[l] = J[r][l] * w[l];
      for (int c = r; c < 9; c++) {
        for (int l = 0; l < 4; l++) pt[l] = pt[l] + Jw[l] * J[c][l];
        pt += 4;
      }
    }
    num += 4; numIn1++;
    shiftUp(false);
  }
  void shiftUp(bool force) {  // :1664-1685
    if (numIn1 > 1000 || force) { for (int i = 0; i < 180; i++) SSEData1k[i] = SSEData[i] + SSEData1k[i]; numIn1k += numIn1; numIn1 = 0; memset(SSEData, 0, sizeof(SSEData)); }
    if (numIn1k > 1000 || force) { for (int i = 0; i < 180; i++) SSEData1m[i] = SSEData1k[i] + SSEData1m[i]; numIn1m += numIn1k; numIn1k = 0; memset(SSEData1k, 0, sizeof(SSEData1k)); }
  }
};

// ScaleAccumulator.h:27-106 : 2x2 upper triangle (J0 = d r/d s, J1 = r), 4 lanes.
struct ScaleAccumulator {
  float hessian[2][2];
  size_t num;
  float sseData[12], sseData1k[12], sseData1m[12];
  float numIn1, numIn1k, numIn1m;
  void initialize() { memset(hessian, 0, sizeof(hessian)); memset(sseData, 0, sizeof(sseData)); memset(sseData1k, 0, sizeof(sseData1k)); memset(sseData1m, 0, sizeof(sseData1m)); num = 0; numIn1 = numIn1k = numIn1m = 0; }
  void finish() {
    memset(hessian, 0, sizeof(hessian));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 2; r++)
      for (int c = r; c < 2; c++) {
        float d = sseData1m[idx + 0] + sseData1m[idx + 1] + sseData1m[idx + 2] + sseData1m[idx + 3];
        hessian[r][c] = hessian[c][r] = d;
        idx += 4;
      }
  }
  void updateSSE_oneed(const float J0[4], const float J1[4], const float w[4]) {  // :60-77
    for (int l = 0; l < 4; l++) {
      float J0w = J0[l] * w[l];
      sseData[0 + l] = sseData[0 + l] + J0w * J0[l];
      sseData[4 + l] = sseData[4 + l] + J0w * J1[l];
      float J1w = J1[l] * w[l];
      sseData[8 + l] = sseData[8 + l] + J1w * J1[l];
    }
    num += 4; numIn1++;
    shiftUp(false);
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) { for (int i = 0; i < 12; i++) sseData1k[i] = sseData[i] + sseData1k[i]; numIn1k += numIn1; numIn1 = 0; memset(sseData, 0, sizeof(sseData)); }
    if (numIn1k > 1000 || force) { for (int i = 0; i < 12; i++) sseData1m[i] = sseData1k[i] + sseData1m[i]; numIn1m += numIn1k; numIn1k = 0; memset(sseData1k, 0, sizeof(sseData1k)); }
  }
};

}  // namespace orc
