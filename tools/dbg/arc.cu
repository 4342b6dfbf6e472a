#include <cstdint>
#include <cstdio>
__device__ __noinline__ int arc_score(const uint8_t* p, int pp) {
  const int v = p[0];
  int d[16];
  d[0] = v - p[3 * pp];      d[1] = v - p[3 * pp + 1];   d[2] = v - p[2 * pp + 2];   d[3] = v - p[pp + 3];
  d[4] = v - p[3];           d[5] = v - p[-pp + 3];      d[6] = v - p[-2 * pp + 2];  d[7] = v - p[-3 * pp + 1];
  d[8] = v - p[-3 * pp];     d[9] = v - p[-3 * pp - 1];  d[10] = v - p[-2 * pp - 2]; d[11] = v - p[-pp - 3];
  d[12] = v - p[-3];         d[13] = v - p[pp - 3];      d[14] = v - p[2 * pp - 2];  d[15] = v - p[3 * pp - 1];
  int mn2[16], mx2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn2[k] = min(d[k], d[(k + 1) & 15]);
    mx2[k] = max(d[k], d[(k + 1) & 15]);
  }
  int mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn4[k] = min(mn2[k], mn2[(k + 2) & 15]);
    mx4[k] = max(mx2[k], mx2[(k + 2) & 15]);
  }
  int best = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
    const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
    best = max(best, max(mn9, -mx9));
  }
  return best;
}
__global__ void k(const uint8_t* p, int pp, int* out){ *out = arc_score(p, pp); }
__host__ __device__ __noinline__ int arc_ref(const uint8_t* p, int pp) {
  const int off[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
  int v = p[0], best = 0;
  for (int k = 0; k < 16; ++k) {
    int mn = 1000, mx = -1000;
    for (int i = 0; i < 9; ++i) { int kk = (k + i) & 15; int d = v - p[off[kk][1] * pp + off[kk][0]]; if(k==9) printf("k9 i=%d kk=%d d=%d\n", i, kk, d); mn = min(mn, d); mx = max(mx, d); }
    best = max(best, max(mn, -mx));
  }
  return best;
}
__global__ void k2(const uint8_t* p, int pp, int* out){ out[0] = arc_score(p, pp); out[1] = arc_ref(p, pp); }
int main(){
  int dv[16]={8,4,0,9,6,3,4,6,1,4,7,5,9,6,3,8};
  const int off[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
  uint8_t h[49]; for(int i=0;i<49;i++) h[i]=132;
  for(int k=0;k<16;k++) h[(3+off[k][1])*7+3+off[k][0]] = 132-dv[k];
  uint8_t* d; int* o; cudaMalloc(&d,49); cudaMalloc(&o,8); cudaMemcpy(d,h,49,cudaMemcpyHostToDevice);
  k2<<<1,1>>>(d+3*7+3,7,o); int r[2]; cudaMemcpy(r,o,8,cudaMemcpyDeviceToHost); printf("host ref=%d\n", arc_ref(h+3*7+3,7)); printf("fast=%d ref=%d err=%s\n", r[0], r[1], cudaGetErrorString(cudaGetLastError()));
}
