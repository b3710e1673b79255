// read_ahead_tsan.cpp -- the read-ahead state machine (pgsd_sph_b200/csrc/read_ahead.cpp) under ThreadSanitizer: 4 reader
// threads (ascending / descending), one thread calling reset() every 0.5 ms, host memory.  Development tool.
//   g++ -O1 -g -std=c++17 -fsanitize=thread -Ipgsd_sph_b200/csrc tools/read_ahead_tsan.cpp pgsd_sph_b200/csrc/read_ahead.cpp -o /tmp/ra_tsan -lpthread && /tmp/ra_tsan
#include "read_ahead.h"
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <thread>
#include <unistd.h>
#include <vector>
using namespace pgsdb;
static bool rd(int fd, void* dst, uint64_t n, uint64_t off){ uint64_t g=0; while(g<n){ ssize_t k=pread(fd,(char*)dst+g,n-g,off+g); if(k<=0) return false; g+=k;} return true; }
static bool al(void** p, uint64_t n){ *p=malloc(n); return *p!=nullptr; }
static void rl(void* p){ free(p); }
static bool cp(void* d, const void* s, uint64_t n){ memcpy(d,s,n); return true; }
int main(){
  const uint64_t CH=384*1024, N=32;
  std::vector<unsigned char> data(CH*N); for(size_t i=0;i<data.size();i++) data[i]=(unsigned char)(i*2654435761u>>13);
  int wfd=open("/tmp/ra_tsan_blob.bin",O_CREAT|O_TRUNC|O_RDWR,0644); if(write(wfd,data.data(),data.size())!=(ssize_t)data.size()) return 2; close(wfd);
  ReadAhead ra(ReadAheadOps{rd,al,rl,cp,nullptr});
  std::atomic<int> bad{0}; std::atomic<bool> stop{false};
  auto reader=[&](int t){ int fd=open("/tmp/ra_tsan_blob.bin",O_RDONLY); std::vector<unsigned char> buf(CH);
    for(int rep=0;rep<8;rep++) for(uint64_t i=0;i<N;i++){ uint64_t k=(t&1)? N-1-i : i; if(!ra.read(fd,buf.data(),CH,k*CH)||memcmp(buf.data(),&data[k*CH],CH)) bad++; }
    close(fd); };
  std::thread r([&]{ while(!stop){ ra.reset(); usleep(500);} });
  std::vector<std::thread> ts; for(int t=0;t<4;t++) ts.emplace_back(reader,t);
  for(auto& t:ts) t.join(); stop=true; r.join();
  uint64_t h,i,d; ra.stats(&h,&i,&d); printf("bad=%d hits=%llu issued=%llu dropped=%llu\n",bad.load(),(unsigned long long)h,(unsigned long long)i,(unsigned long long)d);
  return bad?1:0; }
