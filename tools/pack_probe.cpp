// pack_probe.cpp -- host packing rate of pg_scan pageable path (csrc/host_pack.cpp): g++ -O2 -pthread -o tools/pack_probe tools/pack_probe.cpp pygemma_b200/csrc/host_pack.cpp
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
namespace pg { void pack_block_threads(char* dst, const char* src, size_t pitch, size_t width, size_t rows); }
int main(){
  size_t n=10000, m=100000, mb=25088;
  std::vector<char> X(n*m); for(size_t i=0;i<X.size();i+=4097) X[i]=(char)(i*7);
  char* dst=(char*)aligned_alloc(4096, n*mb); memset(dst,0,n*mb);
  for(int rep=0;rep<3;rep++){
    auto t0=std::chrono::steady_clock::now();
    pg::pack_block_threads(dst, X.data()+ 1024+ (rep? 5120:13), m, mb - (rep==2?3:0), n);
    auto t1=std::chrono::steady_clock::now();
    double ms=std::chrono::duration<double,std::milli>(t1-t0).count();
    size_t w=mb-(rep==2?3:0), off=1024+(rep?5120:13); bool ok=true;
    for(size_t r=0;r<n&&ok;r++) ok = memcmp(dst+r*w, X.data()+off+r*m, w)==0;
    printf("rep %d: %.2f ms  %.1f GB/s ok=%d\n",rep,ms,n*w/ms/1e6,(int)ok);
  }
}
