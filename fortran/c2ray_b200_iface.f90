!> ISO_C_BINDING interfaces to libc2ray_b200.so (include/c2ray_b200.h).
!!
!! This file and evolve_b200.F90 are the reference-side half of the drop-in: together they replace
!! evolve.F90 (module evolve) of C2-Ray3Dm and everything it calls (master_slave.F90,
!! evolve_source.F90, evolve_point.F90, column_density.f90, the look-ups of
!! radiation_photoionrates.F90, doric.f90, and the grid loops of photonstatistics.F90).
!! They cannot be compiled in the development image (it has no Fortran compiler); they are kept
!! mechanical so a maintainer can review them against include/c2ray_b200.h line by line.
module c2ray_b200_iface

  use, intrinsic :: iso_c_binding

  implicit none

  integer(c_int),parameter :: C2B_NUMTAU = 2000
  integer(c_int),parameter :: C2B_MAX_ITER = 104
  integer(c_int),parameter :: C2B_UNIQUE_ID_BYTES = 128

  !> struct c2b_config
  type,bind(C) :: c2b_config
     integer(c_int32_t) :: mesh(3)
     integer(c_int32_t) :: device
     integer(c_int32_t) :: rank, nranks
     integer(c_int32_t) :: isothermal
     integer(c_int32_t) :: type_of_clumping
     integer(c_int32_t) :: use_LLS
     integer(c_int32_t) :: type_of_LLS
     integer(c_int32_t) :: subboxsize
     integer(c_int32_t) :: max_subbox
     integer(c_int32_t) :: max_outer_iter
     integer(c_int32_t) :: reserved0
     real(c_double) :: epsilon
     real(c_double) :: convergence_fraction
     real(c_double) :: minimum_fractional_change
     real(c_double) :: minimum_fraction_of_atoms
     real(c_double) :: loss_fraction
     real(c_double) :: max_coldensh
     real(c_double) :: tau_photo_limit
     real(c_double) :: minlogtau, dlogtau
     real(c_double) :: sigma_HI
     real(c_double) :: pi
     real(c_double) :: sqrt2, sqrt3
     real(c_double) :: bh00, albpow, colh0, temph0
     real(c_double) :: abu_c
     real(c_double) :: k_B
     real(c_double) :: gamma1
     real(c_double) :: minitemp
     real(c_double) :: relative_denergy
     real(c_double) :: tau_heat_limit
     real(c_double) :: H0, Omega0
     integer(c_int32_t) :: cosmological
     integer(c_int32_t) :: reserved1
  end type c2b_config

  !> struct c2b_photon_stats
  type,bind(C) :: c2b_photon_stats
     real(c_double) :: h0_before, h1_before, h0_after, h1_after
     real(c_double) :: totrec, totcollisions, dh0, total_ion
     real(c_double) :: totalsrc, photcons, total_photon_loss, LLS_loss
  end type c2b_photon_stats

  !> struct c2b_pass_report
  type,bind(C) :: c2b_pass_report
     real(c_double) :: photon_loss_all
     integer(c_int64_t) :: sum_nbox_all
     integer(c_int64_t) :: updates
     real(c_double) :: ms_raytrace
     real(c_double) :: ms_allreduce
  end type c2b_pass_report

  !> struct c2b_global_report
  type,bind(C) :: c2b_global_report
     integer(c_int32_t) :: conv_flag
     integer(c_int32_t) :: reserved0
     real(c_double) :: min_avg_neutral
     real(c_double) :: sum_xh_intermed
     type(c2b_photon_stats) :: stats
     real(c_double) :: ms_chemistry
  end type c2b_global_report

  interface

     integer(c_int) function c2b_default_config(cfg) bind(C,name="c2b_default_config")
       import
       type(c2b_config),intent(out) :: cfg
     end function c2b_default_config

     integer(c_int) function c2b_create(cfg,handle) bind(C,name="c2b_create")
       import
       type(c2b_config),intent(in) :: cfg
       type(c_ptr),intent(out) :: handle
     end function c2b_create

     subroutine c2b_destroy(handle) bind(C,name="c2b_destroy")
       import
       type(c_ptr),value :: handle
     end subroutine c2b_destroy

     type(c_ptr) function c2b_last_error(handle) bind(C,name="c2b_last_error")
       import
       type(c_ptr),value :: handle
     end function c2b_last_error

     integer(c_int) function c2b_get_unique_id(id) bind(C,name="c2b_get_unique_id")
       import
       character(kind=c_char),intent(out) :: id(*)
     end function c2b_get_unique_id

     integer(c_int) function c2b_comm_init(handle,id) bind(C,name="c2b_comm_init")
       import
       type(c_ptr),value :: handle
       character(kind=c_char),intent(in) :: id(*)
     end function c2b_comm_init

     integer(c_int) function c2b_set_tables(handle,thick,thin,n) bind(C,name="c2b_set_tables")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: thick(*), thin(*)
       integer(c_int32_t),value :: n
     end function c2b_set_tables

     integer(c_int) function c2b_set_density(handle,ndens) bind(C,name="c2b_set_density")
       import
       type(c_ptr),value :: handle
       real(c_float),intent(in) :: ndens(*)
     end function c2b_set_density

     integer(c_int) function c2b_set_geometry(handle,dr,vol) bind(C,name="c2b_set_geometry")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: dr(3)
       real(c_double),value :: vol
     end function c2b_set_geometry

     integer(c_int) function c2b_cosmo_evol(handle,zfactor) bind(C,name="c2b_cosmo_evol")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: zfactor
     end function c2b_cosmo_evol

     integer(c_int) function c2b_set_clumping_scalar(handle,clumping) bind(C,name="c2b_set_clumping_scalar")
       import
       type(c_ptr),value :: handle
       real(c_float),value :: clumping
     end function c2b_set_clumping_scalar

     integer(c_int) function c2b_set_clumping_grid(handle,grid) bind(C,name="c2b_set_clumping_grid")
       import
       type(c_ptr),value :: handle
       real(c_float),intent(in) :: grid(*)
     end function c2b_set_clumping_grid

     integer(c_int) function c2b_set_lls_scalar(handle,coldensh_LLS) bind(C,name="c2b_set_lls_scalar")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: coldensh_LLS
     end function c2b_set_lls_scalar

     integer(c_int) function c2b_set_lls_grid(handle,grid) bind(C,name="c2b_set_lls_grid")
       import
       type(c_ptr),value :: handle
       real(c_float),intent(in) :: grid(*)
     end function c2b_set_lls_grid

     integer(c_int) function c2b_set_lls_rmax(handle,R_max_LLS) bind(C,name="c2b_set_lls_rmax")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: R_max_LLS
     end function c2b_set_lls_rmax

     integer(c_int) function c2b_set_temperature(handle,temper_val) bind(C,name="c2b_set_temperature")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: temper_val
     end function c2b_set_temperature

     integer(c_int) function c2b_set_sources(handle,NumSrc,srcpos,NormFlux,S_star) bind(C,name="c2b_set_sources")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),value :: NumSrc
       integer(c_int32_t),intent(in) :: srcpos(3,*)   ! srcpos(3,NumSrc), 1-based, as the reference stores it
       real(c_double),intent(in) :: NormFlux(*)       ! NormFlux_stellar(1:NumSrc): pass NormFlux_stellar(1)
       real(c_double),value :: S_star
     end function c2b_set_sources

     integer(c_int) function c2b_set_xh(handle,xh) bind(C,name="c2b_set_xh")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: xh(*)
     end function c2b_set_xh

     integer(c_int) function c2b_begin_step(handle,sum_xh) bind(C,name="c2b_begin_step")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: sum_xh
     end function c2b_begin_step

     integer(c_int) function c2b_pass_all_sources(handle,niter,dt,rep) bind(C,name="c2b_pass_all_sources")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),value :: niter
       real(c_double),value :: dt
       type(c2b_pass_report),intent(out) :: rep
     end function c2b_pass_all_sources

     integer(c_int) function c2b_global_pass(handle,dt,rep) bind(C,name="c2b_global_pass")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: dt
       type(c2b_global_report),intent(out) :: rep
     end function c2b_global_pass

     integer(c_int) function c2b_end_step(handle,dt,converged,final_stats) bind(C,name="c2b_end_step")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: dt
       integer(c_int32_t),value :: converged
       type(c2b_photon_stats),intent(out) :: final_stats
     end function c2b_end_step

     integer(c_int) function c2b_get_xh(handle,xh) bind(C,name="c2b_get_xh")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: xh(*)
     end function c2b_get_xh

     integer(c_int) function c2b_get_xh_av(handle,xh_av) bind(C,name="c2b_get_xh_av")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: xh_av(*)
     end function c2b_get_xh_av

     integer(c_int) function c2b_get_xh_intermed(handle,x) bind(C,name="c2b_get_xh_intermed")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: x(*)
     end function c2b_get_xh_intermed

     integer(c_int) function c2b_get_phih(handle,phih) bind(C,name="c2b_get_phih")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: phih(*)
     end function c2b_get_phih

     integer(c_int) function c2b_set_iter_state(handle,niter,photon_loss_all,phih,xh_av,xh_intermed) &
          bind(C,name="c2b_set_iter_state")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),value :: niter
       real(c_double),value :: photon_loss_all
       real(c_double),intent(in) :: phih(*), xh_av(*), xh_intermed(*)
     end function c2b_set_iter_state

     integer(c_int) function c2b_get_iter_state(handle,niter,photon_loss_all,phih,xh_av,xh_intermed) &
          bind(C,name="c2b_get_iter_state")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),intent(out) :: niter
       real(c_double),intent(out) :: photon_loss_all
       real(c_double),intent(out) :: phih(*), xh_av(*), xh_intermed(*)
     end function c2b_get_iter_state

     integer(c_int) function c2b_device_count() bind(C,name="c2b_device_count")
       import
     end function c2b_device_count

     ! ---- non-isothermal path (isothermal=.false.) ---------------------------------------------------
     integer(c_int) function c2b_set_heat_tables(handle,heat_thick,heat_thin,n) bind(C,name="c2b_set_heat_tables")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: heat_thick(*), heat_thin(*)
       integer(c_int32_t),value :: n
     end function c2b_set_heat_tables

     integer(c_int) function c2b_set_cooling_table(handle,log10_temp,log10_cool,n) bind(C,name="c2b_set_cooling_table")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: log10_temp(*), log10_cool(*)
       integer(c_int32_t),value :: n
     end function c2b_set_cooling_table

     integer(c_int) function c2b_set_redshift(handle,zred) bind(C,name="c2b_set_redshift")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: zred
     end function c2b_set_redshift

     !> temperature_grid(mesh(1),mesh(2),mesh(3)) of type(temperature_states): three default reals per cell;
     !! pass c_loc(temperature_grid) (the derived type is not interoperable by name, its storage is)
     integer(c_int) function c2b_set_temperature_grid(handle,tg) bind(C,name="c2b_set_temperature_grid")
       import
       type(c_ptr),value :: handle
       type(c_ptr),value :: tg
     end function c2b_set_temperature_grid

     integer(c_int) function c2b_get_temperature_grid(handle,tg) bind(C,name="c2b_get_temperature_grid")
       import
       type(c_ptr),value :: handle
       type(c_ptr),value :: tg
     end function c2b_get_temperature_grid

     integer(c_int) function c2b_get_phiheat(handle,phiheat) bind(C,name="c2b_get_phiheat")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(out) :: phiheat(*)
     end function c2b_get_phiheat

     integer(c_int) function c2b_set_iter_state_thermal(handle,phiheat,tg) bind(C,name="c2b_set_iter_state_thermal")
       import
       type(c_ptr),value :: handle
       real(c_double),intent(in) :: phiheat(*)
       type(c_ptr),value :: tg
     end function c2b_set_iter_state_thermal

     !> deterministic_clumping (clumping_module.F90:327-363) evaluated on the device from the resident ndens; p1..p3 =
     !! paramsdcm(1:3) of :350, avg_dens as in :356.  A host that calls this from set_clumping instead of filling
     !! clumping_grid itself saves the upload of the grid with every slice.
     integer(c_int) function c2b_set_clumping_from_density(handle,p1,p2,p3,avg_dens) &
          bind(C,name="c2b_set_clumping_from_density")
       import
       type(c_ptr),value :: handle
       real(c_double),value :: p1,p2,p3,avg_dens
     end function c2b_set_clumping_from_density

     !> per-source subbox counts of the last pass (with npr > 1: of every source, all-reduced), for the MPILOG
     !! diagnostics of do_source (evolve_source.F90:216-219)
     integer(c_int) function c2b_get_source_nbox(handle,nbox) bind(C,name="c2b_get_source_nbox")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),intent(out) :: nbox(*)
     end function c2b_get_source_nbox

     !> which rank traces which source in the next pass (0-based ranks; the library's counterpart of
     !! do_grid_master, master_slave.F90:124-231)
     integer(c_int) function c2b_get_source_owner(handle,owner) bind(C,name="c2b_get_source_owner")
       import
       type(c_ptr),value :: handle
       integer(c_int32_t),intent(out) :: owner(*)
     end function c2b_get_source_owner

     !> sources dealt so far to the work-group shapes: one CTA, one cluster, one warp, handed over by the warp shape
     integer(c_int) function c2b_get_route_counts(handle,counts) bind(C,name="c2b_get_route_counts")
       import
       type(c_ptr),value :: handle
       integer(c_int64_t),intent(out) :: counts(4)
     end function c2b_get_route_counts

  end interface

contains

  !> c2b_last_error as a Fortran string
  function c2b_error_text (handle) result(text)
    type(c_ptr),intent(in) :: handle
    character(len=512) :: text
    character(kind=c_char),pointer :: p(:)
    type(c_ptr) :: cp
    integer :: i
    text=" "
    cp=c2b_last_error(handle)
    if (.not.c_associated(cp)) return
    call c_f_pointer(cp,p,(/512/))
    do i=1,512
       if (p(i) == c_null_char) exit
       text(i:i)=p(i)
    enddo
  end function c2b_error_text

end module c2ray_b200_iface
