// Stand-alone check of fast_arc_score against a host reference on random patches (run on the GPU box).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../rgbd-pl-slam_b200/csrc/fast_score.cuh"
__global__ void k(const uint8_t* p, int n, int* out){ int i = blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i] = plslam::fast_arc_score(p + i*49 + 24, 7); }
static int href(const uint8_t* p, int pp){
  const int off[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
  int v=p[0], best=0;
  for(int k=0;k<16;k++){ int mn=1000,mx=-1000; for(int i=0;i<9;i++){int kk=(k+i)&15; int d=v-p[off[kk][1]*pp+off[kk][0]]; mn=std::min(mn,d); mx=std::max(mx,d);} best=std::max(best,std::max(mn,-mx)); }
  return best;
}
int main(){
  const int n=200000; std::vector<uint8_t> h(n*49); srand(1);
  for(int i=0;i<n;i++){ int base=rand()%256, amp = 1+rand()%128; for(int j=0;j<49;j++){ int v=base+(rand()%(2*amp+1))-amp; h[i*49+j]=(uint8_t)std::min(255,std::max(0,v)); } }
  uint8_t* d; int* o; cudaMalloc(&d,h.size()); cudaMalloc(&o,n*4); cudaMemcpy(d,h.data(),h.size(),cudaMemcpyHostToDevice);
  k<<<(n+255)/256,256>>>(d,n,o); std::vector<int> r(n); cudaMemcpy(r.data(),o,n*4,cudaMemcpyDeviceToHost);
  int bad=0, nz=0; for(int i=0;i<n;i++){ int e=href(&h[i*49+24],7); if(e) nz++; if(e!=r[i]){ if(bad<5) printf("mismatch %d: dev %d host %d\n",i,r[i],e); bad++; } }
  printf("arc_rand: n=%d nonzero=%d mismatches=%d err=%s\n", n, nz, bad, cudaGetErrorString(cudaGetLastError()));
  return bad!=0;
}
