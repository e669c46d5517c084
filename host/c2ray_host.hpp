// C++ host-side mirror of the reference's module interface for the photo-ionization hot path.
//
// The reference is compiled Fortran; its toolchain is absent from the development image, so this is the
// compiled-language host that sits above the C ABI (include/c2ray_b200.h) exactly where
// fortran/evolve_b200.F90 would: module state under the reference's module and variable names, and
// `evolve::evolve3D(time, dt, restart)` with the argument meaning of evolve.F90:83.  It contains no numerics:
// every grid operation happens in libc2ray_b200.so.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace c2ray {

namespace sizes { extern int mesh[3]; }                                   // sizes.f90:33
namespace my_mpi { extern int rank, npr; }                                // mpi.F90 / no_mpi.F90
namespace file_admin {                                                    // file_admin.f90:20-31
extern std::ostream* logf;
extern std::ostream* timefile;
extern std::string dump_dir;   // where iterdump1.bin / iterdump2.bin go
}
namespace grid { extern double dr[3], vol; }                              // grid.F90:25,29
namespace density_module { extern std::vector<float> ndens; }            // density_module.F90:22
namespace ionfractions_module { extern std::vector<double> xh; }         // ionfractions_module.F90:22
namespace temperature_module { extern double temper_val; }               // temperature_module.F90:33
namespace clumping_module { extern float clumping; extern std::vector<float> clumping_grid; }   // :17-18
namespace LLS_module { extern double coldensh_LLS, R_max_LLS; extern std::vector<float> LLS_grid; }  // LLS.F90:79-107
namespace sourceprops {                                                   // sourceprops.F90:56-63
extern int NumSrc;
extern std::vector<int32_t> srcpos;            // srcpos(3,NumSrc), 1-based, column = source
extern std::vector<double> NormFlux_stellar;   // element 0 = source 1
}
namespace radiation_sed_parameters { extern double S_star; }             // radiation_sed_parameters.F90:53
namespace radiation_tables {                                              // radiation_tables.F90:78-79
extern std::vector<double> stellar_photo_thick_table, stellar_photo_thin_table;   // (0:NumTau,1)
}
namespace c2ray_parameters {                                              // c2ray_parameters.f90
extern int type_of_clumping, type_of_LLS;
extern bool use_LLS, isothermal;
extern double convergence_fraction;
}
namespace evolve_data {                                                   // evolve_data.F90:40-60
extern std::vector<double> phih_grid, xh_av, xh_intermed;
extern double photon_loss_all[1];
}
namespace evolve_source { extern int sum_nbox_all; }                     // evolve_source.F90:46
namespace photonstatistics {                                              // photonstatistics.F90:41-55
extern double totrec, totcollisions, dh0, total_ion, LLS_loss, grtotal_ion, grtotal_src, photon_loss[1];
}

namespace evolve {
// subroutine evolve3D (time,dt,restart), evolve.F90:83.  Failures of the device library are reported the
// way the reference reports trouble: a line in logf; `last_error()` holds the text, `ok()` is false.
// restart: 0 = fresh step; 1, 2, 3 = start from iterdump1.bin, iterdump2.bin, iterdump.bin (evolve.F90:349-356)
void evolve3D(double time, double dt, int restart);
// seconds of wall clock between iteration dumps (15 minutes in the reference, evolve.F90:259); a test sets 0 to
// get a dump after every pass_all_sources
extern double dump_interval_seconds;
bool ok();
const std::string& last_error();
int last_niter();
void shutdown();   // releases the device handle (the Fortran program simply exits)
}  // namespace evolve

}  // namespace c2ray
