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
  int reps=5; double t0;
  for(int l=0;l<nlev;++l) if(!m.cells[l].empty()) printf("level %d: leaves %ld ivl %zu rows %zu | ref cells %ld ivl %zu\n", l, (long)m.cells[l].n_cells(), m.cells[l].n_intervals(), m.cells[l].rows(), (long)m.ref[l].n_cells(), m.ref[l].n_intervals());
  t0=now(); for(int r=0;r<reps;++r){ Mesh x; x.init_from_cells(c, CellArray(ca)); } printf("mesh build %.2f ms\n",(now()-t0)/reps*1e3);
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=0;l<nlev;++l){ std::vector<smr_item_fv> it; if(!m.cells[l].empty()) fv_items(m,l,it);} } printf("fv items %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=c.max_level;l>=0;--l){ PhaseItems ph; build_ghost_phase(m,l,ph);} } printf("ghost phases %.2f ms\n",(now()-t0)/reps*1e3); }
  { int l=c.max_level; t0=now(); for(int r=0;r<reps;++r){ PhaseItems ph; build_ghost_phase(m,l,ph);} printf("  ghost phase L %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=1;l<=c.max_level;++l){ LevelSet ps=prediction_set(m,l);} } printf("prediction sets %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=1;l<=c.max_level;++l){ LevelSet ps=prediction_set(m,l); if(ps.empty())continue; locate(ps,m.ref[l]); std::vector<smr_item_pred> items; pred_items(2,1,l,ps,m.ref[l-1],items);} } printf("prediction sets+items %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=3;l<c.max_level;++l){ LevelSet s=detail_set(m,l);} } printf("detail sets %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=3;l<c.max_level;++l){ std::vector<smr_item_detail> items; detail_items(m,l,detail_set(m,l),items);} } printf("detail sets+items %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ for(int l=4;l<=c.max_level;++l){ std::vector<smr_item_tag> items; tag_items(m,l,tag_set(m,l),items);} } printf("tag sets+items %.2f ms\n",(now()-t0)/reps*1e3); }
  { t0=now(); for(int r=0;r<reps;++r){ MeshPlan p; build_plan(m,p);} printf("build_plan total %.2f ms\n",(now()-t0)/reps*1e3); }
  { std::vector<uint8_t> tag(m.nref,1); t0=now(); for(int r=0;r<reps;++r){ CellArray x=cells_from_tags(m,tag.data()); } printf("cells_from_tags %.2f ms\n",(now()-t0)/reps*1e3);
    t0=now(); for(int r=0;r<reps;++r){ CellArray x=m.cells; make_graduation(c,x);} printf("graduation %.2f ms\n",(now()-t0)/reps*1e3); }
  { Mesh m2; m2.init_from_cells(c, CellArray(ca)); t0=now(); for(int r=0;r<reps;++r){ TransferPlan tp; build_transfer(m,m2,tp);} printf("transfer (same mesh) %.2f ms\n",(now()-t0)/reps*1e3); }
}
