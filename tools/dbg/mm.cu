#include <cstdio>
#include <cstdint>
__global__ void k(const int* d, int* out){
  int mn=1000, mx=-1000;
  for(int i=0;i<9;i++){ mn=min(mn,d[i]); mx=max(mx,d[i]); }
  out[0]=mn; out[1]=mx; out[2]=max(mn,-mx); out[3]=max(0,max(mn,-mx));
  int a=d[0],b=d[1],c=d[2];
  out[4]=min(min(a,b),c); out[5]=max(max(a,b),c); out[6]=max(a,max(b,-c)); out[7]=max(a, max(min(b,c), -max(b,c)));
}
int main(){ int h[9]={4,7,5,9,6,3,8,8,4}; int *d,*o; cudaMalloc(&d,36); cudaMalloc(&o,32); cudaMemcpy(d,h,36,cudaMemcpyHostToDevice);
 k<<<1,1>>>(d,o); int r[8]; cudaMemcpy(r,o,32,cudaMemcpyDeviceToHost); for(int i=0;i<8;i++) printf("%d ",r[i]); printf("\n expect 3 9 3 3 4 7 4(max(4,max(7,-5))=7) ...\n"); }
