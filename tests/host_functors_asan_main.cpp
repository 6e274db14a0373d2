// TEST INFRASTRUCTURE: drives tests/host_functors.cu under AddressSanitizer / UBSan on heap buffers of EXACTLY n_blocks * type_size
// bytes, so any functor that reads past the last block of a tensor (on the GPU: past the end of the allocation) is caught.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
extern "C" int hostf_dequant(int, const uint8_t*, long long, void*, int, int);
extern "C" int hostf_fast16(int, const uint8_t*, long long, uint32_t*, int);
int main(){
  int types[] = {2,3,6,7,8,10,11,12,13,14,20,23};
  int ts[] =    {18,20,22,24,34,84,110,144,176,210,18,136};
  int bs[] =    {32,32,32,32,32,256,256,256,256,256,32,256};
  for (int t=0;t<12;++t) for (int n : {1, 3, 16}) {
    size_t bytes=(size_t)n*ts[t];
    uint8_t* exact = new uint8_t[bytes];                            // exact-size, redzone right after
    for (size_t i=0;i<bytes;++i) exact[i]=(uint8_t)(i*131+t);
    std::vector<float> out((size_t)n*bs[t]);
    for (int math=0;math<3;++math) for (int od=0;od<3;++od){
      int rc=hostf_dequant(types[t], exact, n, out.data(), od, math);
      if(rc){printf("rc %d\n",rc);return 1;}
    }
    if (((uintptr_t)exact & 15)==0) { std::vector<uint32_t> o((size_t)n*bs[t]/2); hostf_fast16(types[t], exact, n, o.data(), 1); hostf_fast16(types[t], exact, n, o.data(), 0);} 
    delete[] exact;
  }
  puts("asan run ok");
  return 0;
}
