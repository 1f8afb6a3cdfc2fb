// host_nt_store.cpp — streaming-store (non-temporal) bandwidth of the host with 4/8/12/16 threads and 128/256/512-bit
// stores: the ceiling of the packed read-back's expansion (Engine::copy_out), which writes every raster byte once.
// Build: g++ -O2 -pthread -o host_nt_store host_nt_store.cpp
#include <immintrin.h>
#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <sys/mman.h>
static void fill128(uint8_t* d, size_t n, int v){ __m128i x=_mm_set1_epi8((char)v); __m128i*o=(__m128i*)d; for(size_t i=0;i<n/16;i++) _mm_stream_si128(o+i,x); _mm_sfence(); }
__attribute__((target("avx2"))) static void fill256(uint8_t* d, size_t n, int v){ __m256i x=_mm256_set1_epi8((char)v); __m256i*o=(__m256i*)d; for(size_t i=0;i<n/32;i++) _mm256_stream_si256(o+i,x); _mm_sfence(); }
__attribute__((target("avx512f,avx512bw"))) static void fill512(uint8_t* d, size_t n, int v){ __m512i x=_mm512_set1_epi8((char)v); __m512i*o=(__m512i*)d; for(size_t i=0;i<n/64;i++) _mm512_stream_si512(o+i,x); _mm_sfence(); }
static void fillms(uint8_t* d, size_t n, int v){ memset(d,v,n); }
typedef void (*fn)(uint8_t*,size_t,int);
int main(int argc,char**argv){
  size_t N=(size_t)1<<30; uint8_t* buf=(uint8_t*)mmap(0,N,PROT_READ|PROT_WRITE,MAP_PRIVATE|MAP_ANONYMOUS,-1,0); if(buf==MAP_FAILED){puts("fail");return 1;} madvise(buf,N,MADV_HUGEPAGE); memset(buf,1,N);
  fn fs[4]={fill128,fill256,fill512,fillms}; const char*nm[4]={"nt128","nt256","nt512","memset"};
  int nts[4]={4,8,12,16};
  for(int a=0;a<4;a++) for(int k=0;k<4;k++){ int nt=nts[a];
    double best=1e9;
    for(int rep=0;rep<3;rep++){
      auto t0=std::chrono::steady_clock::now();
      std::vector<std::thread> th; size_t per=N/nt/4096*4096;
      for(int t=0;t<nt;t++) th.emplace_back(fs[k],buf+t*per,per,rep+k);
      for(auto&t:th) t.join();
      double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); if(s<best)best=s;
    }
    printf("%d threads %s: %.1f GB/s\n",nt,nm[k],N/best/1e9); fflush(stdout);
  }
}
