#include <cstdio>
#include <cstdint>
__device__ const int off[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
// A: original style
__device__ __noinline__ int vA(const uint8_t* p, int pp) {
  int v = p[0], best = 0;
  for (int k = 0; k < 16; ++k) {
    int mn = 1000, mx = -1000;
    for (int i = 0; i < 9; ++i) { int kk = (k + i) & 15; int d = v - p[off[kk][1] * pp + off[kk][0]]; mn = min(mn, d); mx = max(mx, d); }
    best = max(best, max(mn, -mx));
  }
  return best;
}
// B: separate polarity passes
__device__ __noinline__ int vB(const uint8_t* p, int pp) {
  int v = p[0], best = 0;
  int q[16];
  for (int k = 0; k < 16; ++k) q[k] = p[off[k][1] * pp + off[k][0]];
  for (int k = 0; k < 16; ++k) {
    int wmx = 0, wmn = 255;
    for (int i = 0; i < 9; ++i) { int kk = (k + i) & 15; wmx = max(wmx, q[kk]); wmn = min(wmn, q[kk]); }
    int a = v - wmx, b = wmn - v;
    if (a > best) best = a;
    if (b > best) best = b;
  }
  return best;
}
// C: only darker polarity (min of d)
__device__ __noinline__ int vC(const uint8_t* p, int pp) {
  int v = p[0], best = -1000;
  for (int k = 0; k < 16; ++k) {
    int mn = 1000;
    for (int i = 0; i < 9; ++i) { int kk = (k + i) & 15; int d = v - p[off[kk][1] * pp + off[kk][0]]; mn = min(mn, d); }
    best = max(best, mn);
  }
  return best;
}
// D: only brighter polarity
__device__ __noinline__ int vD(const uint8_t* p, int pp) {
  int v = p[0], best = -1000;
  for (int k = 0; k < 16; ++k) {
    int mx = -1000;
    for (int i = 0; i < 9; ++i) { int kk = (k + i) & 15; int d = v - p[off[kk][1] * pp + off[kk][0]]; mx = max(mx, d); }
    best = max(best, -mx);
  }
  return best;
}
__global__ void k2(const uint8_t* p, int pp, int* out){ out[0] = vA(p, pp); out[1] = vB(p, pp); out[2]=vC(p,pp); out[3]=vD(p,pp);}
int main(){
  int dv[16]={8,4,0,9,6,3,4,6,1,4,7,5,9,6,3,8};
  const int offh[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
  uint8_t h[49]; for(int i=0;i<49;i++) h[i]=132;
  for(int k=0;k<16;k++) h[(3+offh[k][1])*7+3+offh[k][0]] = 132-dv[k];
  uint8_t* d; int* o; cudaMalloc(&d,49); cudaMalloc(&o,16); cudaMemcpy(d,h,49,cudaMemcpyHostToDevice);
  k2<<<1,1>>>(d+3*7+3,7,o); int r[4]; cudaMemcpy(r,o,16,cudaMemcpyDeviceToHost); printf("A=%d B=%d C=%d D=%d (expect 3 3 3 -4)\n", r[0], r[1], r[2], r[3]);
}
