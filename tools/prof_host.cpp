// Host-side profiling harness: g++ -O2 -std=c++17 -fopenmp -I. -o /tmp/prof_host tools/prof_host.cpp
#include "samurai_b200/csrc/batches.hpp"
#include <chrono>
#include <cstdio>
#include <fstream>
using namespace smr;
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv){
  std::ifstream f(argv[1], std::ios::binary); int64_t n; f.read((char*)&n,8);
  std::vector<int32_t> d(n*5); f.read((char*)d.data(), n*20);
  MeshConfig c; c.dim=2;c.min_level=4;c.max_level=atoi(argv[2]);c.pred_radius=1;
  int nlev=Mesh::levels_for(c); std::vector<SetBuilder> b(nlev);
  for(int64_t i=0;i<n;++i) b[d[5*i]].add(mk_key(d[5*i+1],0), d[5*i+3], d[5*i+4]);
  CellArray ca(nlev); for(int l=0;l<nlev;++l) ca[l]=b[l].build();
  Mesh m; m.init_from_cells(c, CellArray(ca));
  int reps=10; double t0;
  t0=now(); for(int r=0;r<reps;++r){ Mesh x; x.init_from_cells(c, CellArray(ca)); } printf("mesh build %.2f ms\n",(now()-t0)/reps*1e3);
  { MeshPlan p; build_plan(m,p); t0=now(); for(int r=0;r<reps;++r){ build_plan(m,p);} printf("build_plan %.2f ms (arena %.1f MB)\n",(now()-t0)/reps*1e3, p.arena.size/1e6); }
  { std::vector<uint8_t> tag(m.nref,1); t0=now(); for(int r=0;r<reps;++r){ CellArray x=cells_from_tags(m,tag.data()); } printf("cells_from_tags %.2f ms\n",(now()-t0)/reps*1e3);
    t0=now(); for(int r=0;r<reps;++r){ CellArray x=m.cells; make_graduation(c,x);} printf("graduation %.2f ms\n",(now()-t0)/reps*1e3); }
  { Mesh m2; m2.init_from_cells(c, CellArray(ca)); TransferPlan tp; build_transfer(m,m2,tp); t0=now(); for(int r=0;r<reps;++r){ build_transfer(m,m2,tp);} printf("transfer %.2f ms\n",(now()-t0)/reps*1e3); }
}
