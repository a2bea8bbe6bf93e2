// tests/inflate/inflate_sweep.cpp -- TEST ONLY.  Correctness sweep of sph::fast_inflate (secphase_b200/host/sph_inflate.cpp):
// data families x zlib levels 0..9 x strategies (default, filtered, Huffman-only, RLE, fixed) x single/multi-block streams -> raw
// deflate -> the decoder must reproduce the input exactly, refuse a wrong output size, and survive corrupted or truncated
// streams without touching memory outside its buffers (run under ASan/UBSan by tests/test_host_inflate.py).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <random>
#include <zlib.h>
#include "../../secphase_b200/host/sph_bgzf.hpp"
using namespace sph;
int main(){ std::mt19937_64 rng(12345); long cases=0, ok=0, refused=0, wrong=0, fuzz_true=0;
 for(int iter=0; iter<1500; iter++){
   size_t n = (iter%50==0)? 0 : (size_t)(rng()% (iter%7==0? 65536: 20000))+ (iter%3==0?1:0);
   if(n>65535) n=65535;
   std::vector<uint8_t> src(n);
   int fam = iter%6;
   for(size_t i=0;i<n;i++){
     switch(fam){ case 0: src[i]=rng(); break; case 1: src[i]="ACGT"[rng()&3]; break; case 2: src[i]= (i>40 && (rng()%10)) ? src[i-1-(rng()%40)] : (uint8_t)rng(); break;
       case 3: src[i]= 'A'; break; case 4: src[i]= 33 + (rng()%4)*10 + ((rng()%5)==0? rng()%10:0); break; default: src[i]= (uint8_t)(i*7 + (rng()%3)); }
   }
   int level = iter%10; int strat[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
   z_stream zs; memset(&zs,0,sizeof zs); deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strat[(iter/10)%5]);
   std::vector<uint8_t> comp(n + n/100 + 600 + 16); zs.next_in=src.data(); zs.avail_in=n; zs.next_out=comp.data(); zs.avail_out=comp.size()-16;
   // sometimes several deflate blocks per stream
   if(iter%4==0 && n>2000){ zs.avail_in=n/2; deflate(&zs, Z_FULL_FLUSH); zs.avail_in = n - n/2; }
   int rc=deflate(&zs,Z_FINISH); size_t clen=zs.total_out; deflateEnd(&zs); if(rc!=Z_STREAM_END){ printf("deflate failed\n"); return 2; }
   std::vector<uint8_t> exact(comp.begin(), comp.begin()+clen+8);   // exactly 8 readable bytes behind the stream (ASan sees anything beyond)
   std::vector<uint8_t> out(n? n:1, 0xAA);
   bool r = fast_inflate(exact.data(), clen, out.data(), n); cases++;
   if(!r) refused++; else if(n && memcmp(out.data(), src.data(), n)) { wrong++; printf("WRONG iter %d fam %d level %d n %zu\n", iter, fam, level, n);} else ok++;
   // wrong expected size must be refused
   if(n>1 && fast_inflate(exact.data(), clen, out.data(), n-1)) { printf("accepted short output, iter %d\n", iter); wrong++; }
   // corrupted streams: must not crash / write out of bounds (ASan), result may be anything but 'true with wrong size'
   for(int k=0;k<3;k++){ std::vector<uint8_t> bad(exact); if(clen){ bad[rng()%clen] ^= (uint8_t)(1u<<(rng()%8)); if(k==2) bad[rng()%clen]=rng(); }
     size_t cl2 = (k==1 && clen>2) ? clen - 1 - rng()%(clen/2) : clen; std::vector<uint8_t> b2(bad.begin(), bad.begin()+cl2); b2.resize(cl2+8, 0x5A);
     std::vector<uint8_t> o2(n? n:1); if(fast_inflate(b2.data(), cl2, o2.data(), n)) fuzz_true++; }
 }
 printf("cases %ld ok %ld refused %ld wrong %ld (corrupted streams accepted by the decoder, left to the CRC: %ld)\n", cases, ok, refused, wrong, fuzz_true);
 return wrong?1:0; }
