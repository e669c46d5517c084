// Compiled-language driver of the drop-in: plays the part of C2Ray.F90:352-394 for a case file written by
// the tests (tests/test_gpu_host_cpp.py), calling evolve::evolve3D through the host mirror, and writes the
// arrays output.F90 would stream.  Case file (little endian): int32 mesh[3], nsteps, NumSrc,
// type_of_clumping, use_LLS, type_of_LLS; double dt, dr[3], vol, temper_val, clumping, coldensh_LLS,
// R_max_LLS, S_star, zfactor; then ndens f32 N^3, xh f64 N^3, [clumping_grid f32 N^3], [LLS_grid f32 N^3],
// srcpos int32 3*NumSrc, NormFlux f64 NumSrc, thick f64 2001, thin f64 2001.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "c2ray_host.hpp"

using namespace c2ray;

template <class T>
static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: run_case <case.bin> <out.bin> [log.txt [dump_dir [dump_interval_s [restart]]]]\n");
    return 2;
  }
  // optional: where iterdump[12].bin go, the dump interval (0 = after every pass) and the restart flag of the
  // first evolve3D call (1/2: start from iterdump1.bin/iterdump2.bin, evolve.F90:349-356)
  if (argc > 4) file_admin::dump_dir = std::string(argv[4]) + "/";
  if (argc > 5) evolve::dump_interval_seconds = std::atof(argv[5]);
  const int restart_first = (argc > 6) ? std::atoi(argv[6]) : 0;
  std::ifstream f(argv[1], std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  std::ofstream logfile;
  if (argc > 3) { logfile.open(argv[3]); file_admin::logf = &logfile; }
  int32_t hdr[8];
  rd(f, hdr, 8);
  for (int d = 0; d < 3; ++d) sizes::mesh[d] = hdr[d];
  const int nsteps = hdr[3];
  sourceprops::NumSrc = hdr[4];
  c2ray_parameters::type_of_clumping = hdr[5];
  c2ray_parameters::use_LLS = hdr[6] != 0;
  c2ray_parameters::type_of_LLS = hdr[7];
  double sc[13];
  rd(f, sc, 13);
  const double dt = sc[0];
  grid::dr[0] = sc[1]; grid::dr[1] = sc[2]; grid::dr[2] = sc[3];
  grid::vol = sc[4];
  temperature_module::temper_val = sc[5];
  clumping_module::clumping = (float)sc[6];
  LLS_module::coldensh_LLS = sc[7];
  LLS_module::R_max_LLS = sc[8];
  radiation_sed_parameters::S_star = sc[9];
  const double zfactor = sc[10];
  const size_t n = (size_t)sizes::mesh[0] * sizes::mesh[1] * sizes::mesh[2];
  density_module::ndens.resize(n);
  ionfractions_module::xh.resize(n);
  rd(f, density_module::ndens.data(), n);
  rd(f, ionfractions_module::xh.data(), n);
  if (c2ray_parameters::type_of_clumping >= 3) { clumping_module::clumping_grid.resize(n); rd(f, clumping_module::clumping_grid.data(), n); }
  if (c2ray_parameters::use_LLS && c2ray_parameters::type_of_LLS == 2) { LLS_module::LLS_grid.resize(n); rd(f, LLS_module::LLS_grid.data(), n); }
  sourceprops::srcpos.resize(3 * (size_t)sourceprops::NumSrc);
  sourceprops::NormFlux_stellar.resize((size_t)sourceprops::NumSrc);
  rd(f, sourceprops::srcpos.data(), sourceprops::srcpos.size());
  rd(f, sourceprops::NormFlux_stellar.data(), sourceprops::NormFlux_stellar.size());
  radiation_tables::stellar_photo_thick_table.resize(2001);
  radiation_tables::stellar_photo_thin_table.resize(2001);
  rd(f, radiation_tables::stellar_photo_thick_table.data(), 2001);
  rd(f, radiation_tables::stellar_photo_thin_table.data(), 2001);
  if (!f) { std::fprintf(stderr, "short case file\n"); return 2; }

  std::ofstream out(argv[2], std::ios::binary);
  double sim_time = 0.0;
  for (int step = 0; step < nsteps; ++step) {
    // cosmo_evol (cosmology.F90:161-193) on the HOST copies, as C2Ray.F90:367-370 does before every step
    const double z3 = zfactor * zfactor * zfactor;
    for (int d = 0; d < 3; ++d) grid::dr[d] *= zfactor;
    grid::vol *= z3;
    for (auto& v : density_module::ndens) v = (float)((double)v / z3);
    evolve::evolve3D(sim_time, dt, step == 0 ? restart_first : 0);   // C2Ray.F90:379 (iter_restart on the first call)
    if (!evolve::ok()) { std::fprintf(stderr, "evolve3D failed: %s\n", evolve::last_error().c_str()); return 1; }
    sim_time += dt;
    const int32_t niter = evolve::last_niter();
    out.write(reinterpret_cast<const char*>(&niter), sizeof(niter));
    const double st[6] = {photonstatistics::total_ion, photonstatistics::totrec, photonstatistics::totcollisions,
                          photonstatistics::dh0, photonstatistics::grtotal_ion, photonstatistics::grtotal_src};
    out.write(reinterpret_cast<const char*>(st), sizeof(st));
    out.write(reinterpret_cast<const char*>(ionfractions_module::xh.data()), sizeof(double) * n);
    out.write(reinterpret_cast<const char*>(evolve_data::phih_grid.data()), sizeof(double) * n);
  }
  evolve::shutdown();
  return 0;
}
