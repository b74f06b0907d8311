// pack_bench.cpp -- speed and self-check of the host packer (host_pack.hpp) on this machine:
//   g++ -std=c++17 -O2 -Ireadbouncer_b200/csrc tools/pack_bench.cpp readbouncer_b200/csrc/host_pack.o -lpthread -o /tmp/pack_bench
#include "host_pack.hpp"
#include <chrono>
#include <cstdio>
#include <vector>
#include <cstring>
#include <random>
int main(){
  size_t n = 250u*1000*1000 + 17;
  std::vector<uint8_t> b(n); std::mt19937_64 rng(1);
  const char al[]="ACGTacgtNnURYK-\xC1\xD4\x01@[`{"; 
  for(size_t i=0;i<n;++i){ uint64_t r=rng(); b[i] = (r%1000<3)? al[8+(r>>20)%14] : al[(r>>10)%8]; }
  size_t nw=((n+31)/32+15)/16*16; std::vector<uint32_t> lo_(nw+16),hi_(nw+16),bad_(nw+16),lo2(nw),hi2(nw),bad2(nw);
  auto al64=[](std::vector<uint32_t>&v){ return (uint32_t*)(((uintptr_t)v.data()+63)&~(uintptr_t)63); };
  uint32_t *lo=al64(lo_),*hi=al64(hi_),*bad=al64(bad_);   // 64-byte aligned planes, streaming stores: what the library's staging regions are
  // reference
  for(size_t i=0;i<n;++i){ uint32_t c=b[i],u=c&0xDF; bool ok=u=='A'||u=='C'||u=='G'||u=='T'||u=='U';
    if(ok){ lo2[i>>5]|=((c>>1)&1u)<<(i&31); hi2[i>>5]|=((c>>2)&1u)<<(i&31);} else bad2[i>>5]|=1u<<(i&31); }
  printf("isa=%d threads=%d\n",(int)rb::pack_isa(), rb::host_threads());
  for(int rep=0;rep<4;++rep){
    auto t0=std::chrono::steady_clock::now();
    const size_t task=128*1024; size_t nt=(n+task-1)/task;
    rb::parallel_tasks(nt,[&](size_t t){ size_t o=t*task; size_t m=std::min(task,n-o); rb::pack_bases(b.data()+o,m,lo+o/32,hi+o/32,bad+o/32,true);}, nullptr);
    double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    printf("pack %.2f ms  %.1f GB/s\n",s*1e3,n/s/1e9);
  }
  auto t0=std::chrono::steady_clock::now(); rb::pack_bases(b.data(),n,lo,hi,bad,true);
  double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); printf("single %.2f ms %.1f GB/s\n",s*1e3,n/s/1e9);
  printf("equal %d %d %d\n", !memcmp(lo,lo2.data(),(n+31)/32*4), !memcmp(hi,hi2.data(),(n+31)/32*4), !memcmp(bad,bad2.data(),(n+31)/32*4));
}
