!> Replacement for evolve.F90 of C2-Ray3Dm: the same `module evolve` exporting the same
!! `evolve3D(time,dt,restart)` (evolve.F90:61,76,83), with the ray tracing, the rank reduction,
!! the per-cell chemistry and the grid reductions of the photon statistics executed by
!! libc2ray_b200.so on the GPU.  The outer convergence loop, its log lines and the iteration
!! dumps stay here, written exactly as in the reference, so results/C2Ray.log and
!! results/Timings.log keep their format.
!!
!! Build: in makefile_core replace `evolve.o` by `c2ray_b200_iface.o evolve_b200.o` in the
!! EVOLVE= line (makefile_core:31) and add `-L<repo>/c2ray3dm_b200 -lc2ray_b200` to the link
!! line; master_slave.o, evolve_source.o, evolve_point.o and column_density.o are no longer needed
!! by `evolve` (photonstatistics.o still provides the module variables output.F90 prints).
!!
!! NOT compiled in the development image (no Fortran compiler there): see INTEGRATION.md.
module evolve

  use, intrinsic :: iso_c_binding
  use precision, only: dp
  use my_mpi ! rank, npr, MPI_COMM_NEW
  use file_admin, only: logf, timefile, iterdump, results_dir, dump_dir
  use clocks, only: timestamp_wallclock
  use sizes, only: Ndim, mesh
  use grid, only: dr, vol
  use density_module, only: ndens
  use ionfractions_module, only: xh
  use temperature_module, only: temper_val
  use clumping_module, only: clumping, clumping_grid
  use LLS_module, only: coldensh_LLS, LLS_grid, R_max_LLS
  use sourceprops, only: NumSrc, srcpos, NormFlux_stellar
  use radiation_sed_parameters, only: S_star
  use radiation_sizes, only: NumTau, NumFreqBnd
  use radiation_tables, only: stellar_photo_thick_table, stellar_photo_thin_table
  use c2ray_parameters, only: convergence_fraction, isothermal, use_LLS, type_of_LLS, &
       type_of_clumping, subboxsize, max_subbox, loss_fraction, epsilon, &
       minimum_fractional_change, minimum_fraction_of_atoms
  use photonstatistics, only: photon_loss, LLS_loss, totrec, totcollisions, dh0, total_ion, &
       grtotal_ion, grtotal_src
  use evolve_data, only: phih_grid, xh_av, xh_intermed, photon_loss_all
  use c2ray_b200_iface

  implicit none

  save

  private

  public :: evolve3D

  type(c_ptr) :: handle = c_null_ptr
  integer :: sum_nbox_all

  real(kind=dp) :: sum_xh1_int, sum_xh0_int
  real(kind=dp) :: prev_sum_xh1_int, prev_sum_xh0_int
  real(kind=dp) :: rel_change_sum_xh1, rel_change_sum_xh0

contains

  !> creates the device handle on first use (replaces the allocations of evolve_ini,
  !! evolve_data.F90:73-93, which are still needed for the host copies output.F90 reads)
  subroutine b200_init ()

    type(c2b_config) :: cfg
    character(kind=c_char) :: id(C2B_UNIQUE_ID_BYTES)
    integer :: ierr
#ifdef MPI
    integer :: mympierror
#endif

    ierr=c2b_default_config(cfg)
    cfg%mesh(:)=mesh(:)
    cfg%rank=rank
    cfg%nranks=npr
    cfg%device=mod(rank,8)            ! one rank per GPU of the node
    cfg%type_of_clumping=type_of_clumping
    cfg%use_LLS=merge(1,0,use_LLS)
    cfg%type_of_LLS=type_of_LLS
    cfg%subboxsize=subboxsize
    cfg%max_subbox=max_subbox
    cfg%loss_fraction=loss_fraction
    cfg%epsilon=epsilon
    cfg%convergence_fraction=convergence_fraction
    cfg%minimum_fractional_change=minimum_fractional_change
    cfg%minimum_fraction_of_atoms=minimum_fraction_of_atoms
    cfg%isothermal=merge(1,0,isothermal)
    ierr=c2b_create(cfg,handle)
    if (ierr /= 0) then
       write(logf,*) "c2b_create failed with code ",ierr
       stop "c2ray_b200: cannot create the device handle"
    endif
    if (npr > 1) then
       if (rank == 0) ierr=c2b_get_unique_id(id)
#ifdef MPI
       call MPI_BCAST(id,C2B_UNIQUE_ID_BYTES,MPI_CHARACTER,0,MPI_COMM_NEW,mympierror)
#endif
       ierr=c2b_comm_init(handle,id)
    endif
    ! rad_ini (radiation_tables.F90:95) has already run on the host: hand over its tables
    ierr=c2b_set_tables(handle,stellar_photo_thick_table(0:NumTau,1), &
         stellar_photo_thin_table(0:NumTau,1),NumTau+1)

  end subroutine b200_init

  !> marshals the module state evolve3D reads (SURVEY 8b "hidden inputs")
  subroutine b200_upload_state ()

    integer :: ierr

    ierr=c2b_set_density(handle,ndens)          ! changes every step: cosmo_evol (cosmology.F90:186)
    ierr=c2b_set_geometry(handle,dr,vol)        ! cosmology.F90:181-183
    ierr=c2b_set_temperature(handle,temper_val)
    if (type_of_clumping >= 3) then
       ierr=c2b_set_clumping_grid(handle,clumping_grid)
    else
       ierr=c2b_set_clumping_scalar(handle,clumping)
    endif
    if (use_LLS) then
       select case (type_of_LLS)
       case(1)
          ierr=c2b_set_lls_scalar(handle,coldensh_LLS)
       case(2)
          ierr=c2b_set_lls_grid(handle,LLS_grid)
       case(3)
          ierr=c2b_set_lls_rmax(handle,R_max_LLS)
       end select
    endif
    ierr=c2b_set_sources(handle,NumSrc,srcpos,NormFlux_stellar(1:NumSrc),S_star)
    ierr=c2b_set_xh(handle,xh)

  end subroutine b200_upload_state

  !> evolve.F90:83-281
  subroutine evolve3D (time,dt,restart)

    real(kind=dp),intent(in) :: time !< time
    real(kind=dp),intent(in) :: dt !< time step
    integer,intent(in) :: restart !< restart flag

    integer :: niter
    integer :: conv_flag
    integer :: conv_criterion
    integer :: ierr
    integer(c_int32_t) :: converged
    type(c2b_pass_report) :: pass_rep
    type(c2b_global_report) :: glob_rep
    type(c2b_photon_stats) :: stats

    if (.not.c_associated(handle)) call b200_init ()
    call b200_upload_state ()

    if (restart == 0) then
       ! state_before ; xh_av=xh ; xh_intermed=xh  (evolve.F90:136-147)
       ierr=c2b_begin_step(handle,sum_xh1_int)
       niter=0
       conv_flag=mesh(1)*mesh(2)*mesh(3)
       prev_sum_xh1_int=2.0*mesh(1)*mesh(2)*mesh(3)
       prev_sum_xh0_int=2.0*mesh(1)*mesh(2)*mesh(3)
       rel_change_sum_xh1=1.0
       rel_change_sum_xh0=1.0
    else
       ! start_from_dump (evolve.F90:328-426) reads niter, photon_loss_all, phih_grid, xh_av,
       ! xh_intermed from iterdump[12].bin into the host arrays exactly as before; then:
       ierr=c2b_begin_step(handle,sum_xh1_int)
       ierr=c2b_set_iter_state(handle,niter,photon_loss_all(1),phih_grid,xh_av,xh_intermed)
       ierr=c2b_global_pass(handle,dt,glob_rep)
       conv_flag=glob_rep%conv_flag
       sum_xh1_int=glob_rep%sum_xh_intermed
    endif

    conv_criterion=min(int(convergence_fraction*mesh(1)*mesh(2)*mesh(3)),(NumSrc-1)/3)

    if (rank == 0) write(timefile,"(A,F8.1)") &
         "Time before starting iteration: ", timestamp_wallclock ()

    converged=0
    do
       ! sum_xh1_int=sum(xh_intermed) arrives from the device (fused into the per-cell kernel)
       sum_xh0_int=real(mesh(1)*mesh(2)*mesh(3)) - sum_xh1_int
       if (sum_xh1_int > 0.0) then
          rel_change_sum_xh1=abs(sum_xh1_int-prev_sum_xh1_int)/sum_xh1_int
       else
          rel_change_sum_xh1=1.0
       endif
       if (sum_xh0_int > 0.0) then
          rel_change_sum_xh0=abs(sum_xh0_int-prev_sum_xh0_int)/sum_xh0_int
       else
          rel_change_sum_xh0=1.0
       endif
       if (rank == 0) then
          write(logf,*) "Convergence tests: "
          write(logf,*) "   Test 1 values: ",conv_flag, conv_criterion
          write(logf,*) "   Test 2 values: ",rel_change_sum_xh1, &
               rel_change_sum_xh0, convergence_fraction
       endif
       if (conv_flag < conv_criterion .or. &
            ( rel_change_sum_xh1 < convergence_fraction .and. &
            rel_change_sum_xh0 < convergence_fraction )) then
          converged=1
          if (rank == 0) write(logf,*) "Multiple sources convergence reached"
          exit
       else
          if (niter > 100) then
             if (rank == 0) write(logf,*) 'Multiple sources not converging'
             exit
          endif
       endif
       prev_sum_xh1_int=sum_xh1_int
       prev_sum_xh0_int=sum_xh0_int
       niter=niter+1

       ! set_rates_to_zero + pass_all_sources (+ the all-reduces of evolve.F90:577-616)
       if (rank == 0) write(logf,*) 'Doing all sources '
       ierr=c2b_pass_all_sources(handle,niter,dt,pass_rep)
       photon_loss_all(1)=pass_rep%photon_loss_all
       sum_nbox_all=int(pass_rep%sum_nbox_all)
       if (rank == 0) &
            write(logf,*) "Average number of subboxes: ", &
            real(sum_nbox_all)/real(NumSrc)

       ! global_pass (evolve.F90:499-573)
       ierr=c2b_global_pass(handle,dt,glob_rep)
       conv_flag=glob_rep%conv_flag
       sum_xh1_int=glob_rep%sum_xh_intermed
       if (rank == 0) then
          write(logf,*) "min value avg neutral fraction: ",glob_rep%min_avg_neutral
          write(logf,*) 'Doing global '
          write(logf,*) "Number of non-converged points: ",conv_flag
          write(logf,*) "Intermediate result for mean H ionization fraction: ", &
               sum_xh1_int/real(mesh(1)*mesh(2)*mesh(3))
       endif
       call absorb_stats (glob_rep%stats)
       call report_stats (glob_rep%stats)

       if (rank == 0) write(timefile,"(A,I3,A,F8.1)") &
            "Time after iteration ",niter," : ", timestamp_wallclock ()
    enddo

    ! xh=xh_intermed if converged ; calculate_photon_statistics(dt,xh,xh_av) ; grand totals
    ierr=c2b_end_step(handle,dt,converged,stats)
    call absorb_stats (stats)
    call report_stats (stats)
    grtotal_src=grtotal_src+stats%totalsrc
    grtotal_ion=grtotal_ion+total_ion-totcollisions

    ! host copies for output.F90 (streams 2 and 3) and for the next call
    ierr=c2b_get_xh(handle,xh)
    ierr=c2b_get_phih(handle,phih_grid)

  end subroutine evolve3D

  !> copies the device-side statistics into the photonstatistics module variables output.F90 prints
  subroutine absorb_stats (s)
    type(c2b_photon_stats),intent(in) :: s
    totrec=s%totrec
    totcollisions=s%totcollisions
    dh0=s%dh0
    total_ion=s%total_ion
    LLS_loss=s%LLS_loss
    photon_loss(1)=photon_loss_all(1)/(real(mesh(1))*real(mesh(2))*real(mesh(3)))
  end subroutine absorb_stats

  !> report_photonstatistics, photonstatistics.F90:254-281
  subroutine report_stats (s)
    type(c2b_photon_stats),intent(in) :: s
    if (rank == 0) then
       write(logf,"(8(1pe10.3))") &
            s%total_ion, s%totalsrc, s%photcons, s%dh0/s%total_ion, &
            s%totrec/s%total_ion, s%LLS_loss/s%totalsrc, &
            s%total_photon_loss/s%totalsrc, s%totcollisions/s%total_ion
       write(logf,*) s%h1_before,s%h1_after
    endif
  end subroutine report_stats

end module evolve
