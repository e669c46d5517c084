!> Replacement for evolve.F90 of C2-Ray3Dm: the same `module evolve` exporting the same
!! `evolve3D(time,dt,restart)` (evolve.F90:61,76,83), with the ray tracing, the rank reduction,
!! the per-cell chemistry (and thermal evolution) and the grid reductions of the photon statistics
!! executed by libc2ray_b200.so on the GPU.  The outer convergence loop, its log lines, the
!! 15-minute iteration dumps and the restart from them stay here, written as in the reference, so
!! results/C2Ray.log, results/Timings.log and iterdump[12].bin keep their format.
!!
!! Build: in makefile_core replace `evolve.o` by `c2ray_b200_iface.o evolve_b200.o` in the
!! EVOLVE= line (makefile_core:31) and add `-L<repo>/c2ray3dm_b200 -lc2ray_b200` to the link
!! line; master_slave.o, evolve_source.o, evolve_point.o and column_density.o are no longer needed
!! by `evolve` (photonstatistics.o still provides the module variables output.F90 prints).
!!
!! NOT compiled in the development image (no Fortran compiler there): see INTEGRATION.md.  The
!! C++ twin host/evolve.cpp is the same sequence of calls and IS compiled and tested.
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
  use temperature_module, only: temper_val, temperature_grid
  use clumping_module, only: clumping, clumping_grid
  use LLS_module, only: coldensh_LLS, LLS_grid, R_max_LLS
  use sourceprops, only: NumSrc, srcpos, NormFlux_stellar
  use radiation_sed_parameters, only: S_star
  use radiation_sizes, only: NumTau, NumFreqBnd
  use radiation_tables, only: stellar_photo_thick_table, stellar_photo_thin_table, &
       stellar_heat_thick_table, stellar_heat_thin_table
  use cosmology, only: zred
  use c2ray_parameters, only: convergence_fraction, isothermal, use_LLS, type_of_LLS, &
       type_of_clumping, subboxsize, max_subbox, loss_fraction, epsilon, &
       minimum_fractional_change, minimum_fraction_of_atoms, cosmological, minitemp, &
       relative_denergy
  use photonstatistics, only: photon_loss, LLS_loss, totrec, totcollisions, dh0, total_ion, &
       grtotal_ion, grtotal_src
  use evolve_data, only: phih_grid, phiheat_grid, xh_av, xh_intermed, photon_loss_all
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

  ! what the device already holds (uploads happen only when the host changed the data)
  integer :: dev_NumSrc = -1
  integer,allocatable :: dev_srcpos(:,:)
  real(kind=dp),allocatable :: dev_normflux(:)
  real(kind=dp) :: dev_clumping_sig = -1.0_dp, dev_LLS_sig = -1.0_dp

#ifdef MPI
  integer :: mympierror
#endif

contains

  !> every call of the library is checked: on failure the text of c2b_last_error goes to the log and
  !! the run stops on ALL ranks (a rank that stays behind would hang in the NCCL all-reduce)
  subroutine check (ierr,what)
    integer,intent(in) :: ierr
    character(len=*),intent(in) :: what
    if (ierr /= 0) then
       write(logf,*) "c2ray_b200: ",what," failed with code ",ierr,": ", &
            trim(c2b_error_text(handle))
       flush(logf)
#ifdef MPI
       call MPI_ABORT(MPI_COMM_NEW,ierr,mympierror)
#endif
       stop "c2ray_b200: device library error"
    endif
  end subroutine check

  !> creates the device handle on first use (replaces the allocations of evolve_ini,
  !! evolve_data.F90:73-93, which are still needed for the host copies output.F90 reads)
  subroutine b200_init ()

    type(c2b_config) :: cfg
    character(kind=c_char) :: id(C2B_UNIQUE_ID_BYTES)
    integer :: ierr, local_rank, ndev
#ifdef MPI
    integer :: node_comm
#endif

    call check(c2b_default_config(cfg),"c2b_default_config")
    cfg%mesh(:)=mesh(:)
    cfg%rank=rank
    cfg%nranks=npr
    ! one rank per GPU: the device ordinal is the rank WITHIN the node
    local_rank=0
#ifdef MPI
    call MPI_COMM_SPLIT_TYPE(MPI_COMM_NEW,MPI_COMM_TYPE_SHARED,rank,MPI_INFO_NULL, &
         node_comm,mympierror)
    call MPI_COMM_RANK(node_comm,local_rank,mympierror)
    call MPI_COMM_FREE(node_comm,mympierror)
#endif
    ndev=c2b_device_count()
    if (ndev < 1) call check(110,"c2b_device_count (no CUDA device; there is no CPU fallback)")
    cfg%device=mod(local_rank,ndev)
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
    cfg%cosmological=merge(1,0,cosmological)
    cfg%minitemp=minitemp
    cfg%relative_denergy=relative_denergy
    ierr=c2b_create(cfg,handle)
    if (ierr /= 0) then
       write(logf,*) "c2b_create failed with code ",ierr,": ",trim(c2b_error_text(c_null_ptr))
       stop "c2ray_b200: cannot create the device handle"
    endif
    if (npr > 1) then
       if (rank == 0) call check(c2b_get_unique_id(id),"c2b_get_unique_id")
#ifdef MPI
       call MPI_BCAST(id,C2B_UNIQUE_ID_BYTES,MPI_CHARACTER,0,MPI_COMM_NEW,mympierror)
#endif
       call check(c2b_comm_init(handle,id),"c2b_comm_init")
    endif
    ! rad_ini (radiation_tables.F90:95) has already run on the host: hand over its tables
    call check(c2b_set_tables(handle,stellar_photo_thick_table(0:NumTau,1), &
         stellar_photo_thin_table(0:NumTau,1),NumTau+1),"c2b_set_tables")
    if (.not.isothermal) then
       call check(c2b_set_heat_tables(handle,stellar_heat_thick_table(0:NumTau,1), &
            stellar_heat_thin_table(0:NumTau,1),NumTau+1),"c2b_set_heat_tables")
       call b200_upload_cooling_table ()
    endif

  end subroutine b200_init

  !> the 61 rows of tables/corocool.tab, read as setup_cool does (cooling.f90:64-87): the module
  !! radiative_cooling keeps its table private, so the shim reads the file itself
  subroutine b200_upload_cooling_table ()
    real(kind=dp) :: temp(61), cool(61)
    integer :: itemp
    open(unit=22,file='tables/corocool.tab',status='old')
    do itemp=1,61
       read(22,*) temp(itemp),cool(itemp)
    enddo
    close(22)
    call check(c2b_set_cooling_table(handle,temp,cool,61),"c2b_set_cooling_table")
  end subroutine b200_upload_cooling_table

  !> a cheap signature of a grid the host may have replaced (a new density slice brings new clumping
  !! / LLS grids): size plus a strided sample
  function grid_signature (a) result(sig)
    real,intent(in) :: a(:,:,:)
    real(kind=dp) :: sig
    sig=real(size(a),dp)+sum(real(a(1:size(a,1):7,1:size(a,2):7,1:size(a,3):7),dp))
  end function grid_signature

  !> marshals the module state evolve3D reads (SURVEY 8b "hidden inputs").  ndens, dr, vol and xh
  !! change every step (cosmology.F90:181-186, C2Ray.F90:367-379); sources, clumping and LLS grids
  !! only with a new slice, so they are uploaded only when they differ from what the device holds
  !! (re-sending the sources would also discard the per-source trace lengths the library keeps
  !! for its work queue).
  subroutine b200_upload_state ()

    logical :: new_sources
    real(kind=dp) :: sig

    call check(c2b_set_density(handle,ndens),"c2b_set_density")
    call check(c2b_set_geometry(handle,dr,vol),"c2b_set_geometry")
    if (isothermal) then
       call check(c2b_set_temperature(handle,temper_val),"c2b_set_temperature")
    else
       call check(c2b_set_temperature_grid(handle,c_loc(temperature_grid)),"c2b_set_temperature_grid")
       call check(c2b_set_redshift(handle,zred),"c2b_set_redshift")
    endif
    if (type_of_clumping >= 3) then
       sig=grid_signature(clumping_grid)
       if (sig /= dev_clumping_sig) then
          call check(c2b_set_clumping_grid(handle,clumping_grid),"c2b_set_clumping_grid")
          dev_clumping_sig=sig
       endif
    else
       call check(c2b_set_clumping_scalar(handle,clumping),"c2b_set_clumping_scalar")
    endif
    if (use_LLS) then
       select case (type_of_LLS)
       case(1)
          call check(c2b_set_lls_scalar(handle,coldensh_LLS),"c2b_set_lls_scalar")
       case(2)
          sig=grid_signature(LLS_grid)
          if (sig /= dev_LLS_sig) then
             call check(c2b_set_lls_grid(handle,LLS_grid),"c2b_set_lls_grid")
             dev_LLS_sig=sig
          endif
       case(3)
          call check(c2b_set_lls_rmax(handle,R_max_LLS),"c2b_set_lls_rmax")
       end select
    endif
    new_sources=(NumSrc /= dev_NumSrc)
    if (.not.new_sources .and. NumSrc > 0) then
       new_sources=any(srcpos(:,1:NumSrc) /= dev_srcpos(:,:)) .or. &
            any(NormFlux_stellar(1:NumSrc) /= dev_normflux(:))
    endif
    if (new_sources) then
       call check(c2b_set_sources(handle,NumSrc,srcpos,NormFlux_stellar(1:NumSrc),S_star), &
            "c2b_set_sources")
       if (allocated(dev_srcpos)) deallocate(dev_srcpos,dev_normflux)
       allocate(dev_srcpos(3,NumSrc),dev_normflux(NumSrc))
       dev_srcpos(:,:)=srcpos(:,1:NumSrc)
       dev_normflux(:)=NormFlux_stellar(1:NumSrc)
       dev_NumSrc=NumSrc
    endif
    call check(c2b_set_xh(handle,xh),"c2b_set_xh")

  end subroutine b200_upload_state

  !> write_iteration_dump, evolve.F90:285-324: same files, same records.  The arrays are fetched from
  !! the device at the reference's dump point, between pass_all_sources and global_pass.
  subroutine write_iteration_dump (niter)

    integer,intent(in) :: niter  ! iteration counter
    integer :: ndump=0
    integer(c_int32_t) :: niter_dev
    character(len=20) :: iterfile

    write(timefile,"(A,F8.1)") &
         "Time before writing iterdump: ", timestamp_wallclock ()

    call check(c2b_get_iter_state(handle,niter_dev,photon_loss_all(1),phih_grid,xh_av,xh_intermed), &
         "c2b_get_iter_state")
    if (.not.isothermal) then
       call check(c2b_get_phiheat(handle,phiheat_grid),"c2b_get_phiheat")
       call check(c2b_get_temperature_grid(handle,c_loc(temperature_grid)),"c2b_get_temperature_grid")
    endif

    ndump=ndump+1
    if (mod(ndump,2) == 0) then
       iterfile="iterdump2.bin"
    else
       iterfile="iterdump1.bin"
    endif

    open(unit=iterdump,file=trim(adjustl(dump_dir))//iterfile,form="unformatted", &
         status="unknown")

    write(iterdump) niter
    write(iterdump) photon_loss_all
    write(iterdump) phih_grid
    write(iterdump) xh_av
    write(iterdump) xh_intermed
    if (.not.isothermal) then
       write(iterdump) phiheat_grid
       write(iterdump) temperature_grid
    endif
    close(iterdump)

    write(timefile,"(A,F8.1)") &
         "Time after writing iterdump: ", timestamp_wallclock ()

  end subroutine write_iteration_dump

  !> start_from_dump, evolve.F90:328-426: rank 0 reads the file, everybody receives the records, the
  !! device gets them through c2b_set_iter_state
  subroutine start_from_dump (restart,niter)

    integer,intent(in) :: restart  ! restart flag
    integer,intent(out) :: niter  ! iteration counter
    character(len=20) :: iterfile

    niter=0
    if (restart == 0) then
       if (rank == 0) &
            write(logf,*) "Warning: start_from_dump called incorrectly"
    else
       if (rank == 0) then
          write(timefile,"(A,F8.1)") &
               "Time before reading iterdump: ", timestamp_wallclock ()
          select case (restart)
          case (1)
             iterfile="iterdump1.bin"
          case (2)
             iterfile="iterdump2.bin"
          case (3)
             iterfile="iterdump.bin"
          end select
          open(unit=iterdump,file=trim(adjustl(dump_dir))//iterfile, &
               form="unformatted",status="old")
          read(iterdump) niter
          read(iterdump) photon_loss_all
          read(iterdump) phih_grid
          read(iterdump) xh_av
          read(iterdump) xh_intermed
          if (.not.isothermal) then
             read(iterdump) phiheat_grid
             read(iterdump) temperature_grid
          endif
          close(iterdump)
          write(logf,*) "Read iteration ",niter," from dump file"
          write(logf,*) 'photon loss counter: ',photon_loss_all
          write(logf,*) "Intermediate result for mean ionization fraction: ", &
               sum(xh_intermed(:,:,:))/real(mesh(1)*mesh(2)*mesh(3))
       endif
#ifdef MPI
       call MPI_BCAST(niter,1,MPI_INTEGER,0,MPI_COMM_NEW,mympierror)
       call MPI_BCAST(photon_loss_all,NumFreqBnd,MPI_DOUBLE_PRECISION,0,MPI_COMM_NEW,mympierror)
       call MPI_BCAST(phih_grid,mesh(1)*mesh(2)*mesh(3),MPI_DOUBLE_PRECISION,0,MPI_COMM_NEW,mympierror)
       call MPI_BCAST(xh_av,mesh(1)*mesh(2)*mesh(3),MPI_DOUBLE_PRECISION,0,MPI_COMM_NEW,mympierror)
       call MPI_BCAST(xh_intermed,mesh(1)*mesh(2)*mesh(3),MPI_DOUBLE_PRECISION,0,MPI_COMM_NEW,mympierror)
       if (.not.isothermal) then
          call MPI_BCAST(phiheat_grid,mesh(1)*mesh(2)*mesh(3),MPI_DOUBLE_PRECISION,0,MPI_COMM_NEW,mympierror)
          call MPI_BCAST(temperature_grid,mesh(1)*mesh(2)*mesh(3)*3,MPI_REAL,0,MPI_COMM_NEW,mympierror)
       endif
#endif
       call check(c2b_set_iter_state(handle,niter,photon_loss_all(1),phih_grid,xh_av,xh_intermed), &
            "c2b_set_iter_state")
       if (.not.isothermal) call check(c2b_set_iter_state_thermal(handle,phiheat_grid, &
            c_loc(temperature_grid)),"c2b_set_iter_state_thermal")
       write(timefile,"(A,F8.1)") &
            "Time after reading iterdump: ", timestamp_wallclock ()
    endif

  end subroutine start_from_dump

  !> evolve.F90:83-281
  subroutine evolve3D (time,dt,restart)

    real(kind=dp),intent(in) :: time !< time
    real(kind=dp),intent(in) :: dt !< time step
    integer,intent(in) :: restart !< restart flag

    integer :: niter
    integer :: conv_flag
    integer :: conv_criterion
    integer(c_int32_t) :: converged
    integer(kind=8) :: wallclock1, wallclock2, countspersec
    type(c2b_pass_report) :: pass_rep
    type(c2b_global_report) :: glob_rep
    type(c2b_photon_stats) :: stats

    ! Initialize wall clock counter (for dumps)
    call system_clock(wallclock1)

    if (.not.c_associated(handle)) call b200_init ()
    call b200_upload_state ()

    ! state_before(xh) ; xh_av=xh ; xh_intermed=xh  (evolve.F90:136-147; on a restart the dump
    ! overwrites the two work arrays right below, as in the reference)
    call check(c2b_begin_step(handle,sum_xh1_int),"c2b_begin_step")
    if (restart == 0) then
       niter=0
       conv_flag=mesh(1)*mesh(2)*mesh(3)
       prev_sum_xh1_int=2.0*mesh(1)*mesh(2)*mesh(3)
       prev_sum_xh0_int=2.0*mesh(1)*mesh(2)*mesh(3)
       rel_change_sum_xh1=1.0
       rel_change_sum_xh0=1.0
    else
       ! Reload xh_av,xh_intermed,photon_loss,niter ; global_pass (evolve.F90:154-158)
       call start_from_dump(restart,niter)
       call check(c2b_global_pass(handle,dt,glob_rep),"c2b_global_pass")
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
       call check(c2b_pass_all_sources(handle,niter,dt,pass_rep),"c2b_pass_all_sources")
       photon_loss_all(1)=pass_rep%photon_loss_all
       sum_nbox_all=int(pass_rep%sum_nbox_all)
       if (rank == 0) &
            write(logf,*) "Average number of subboxes: ", &
            real(sum_nbox_all)/real(NumSrc)

       if (rank == 0) then
          call system_clock(wallclock2,countspersec)
          ! Write iteration dump if more than 15 minutes have passed (evolve.F90:248-266)
          write(logf,*) "Time and limit are: ", &
               wallclock2-wallclock1, 15.0*60.0*countspersec
          if (wallclock2-wallclock1 > 15*60*countspersec .or. &
               wallclock2-wallclock1 < 0 ) then
             call write_iteration_dump(niter)
             wallclock1=wallclock2
          endif
       endif

       ! global_pass (evolve.F90:499-573)
       call check(c2b_global_pass(handle,dt,glob_rep),"c2b_global_pass")
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

    ! xh=xh_intermed (and the final temperature) if converged ; calculate_photon_statistics(dt,xh,xh_av)
    call check(c2b_end_step(handle,dt,converged,stats),"c2b_end_step")
    call absorb_stats (stats)
    call report_stats (stats)
    grtotal_src=grtotal_src+stats%totalsrc
    grtotal_ion=grtotal_ion+total_ion-totcollisions

    ! host copies for output.F90 (streams 2 and 3) and for the next call
    call check(c2b_get_xh(handle,xh),"c2b_get_xh")
    call check(c2b_get_phih(handle,phih_grid),"c2b_get_phih")
    if (.not.isothermal) then
       call check(c2b_get_phiheat(handle,phiheat_grid),"c2b_get_phiheat")
       call check(c2b_get_temperature_grid(handle,c_loc(temperature_grid)),"c2b_get_temperature_grid")
    endif

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
