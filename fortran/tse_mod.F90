!> tse_mod: ISO_C_BINDING shim between transport_se's Fortran host and the B200 tracer-advection library (include/tse.h).
!!
!! Drop-in replacement for src/share/cuda_mod.F90: it exports the SAME six entry points the reference calls under
!! `#if USE_CUDA_FORTRAN` (cuda_mod.F90:168,429,451,473,601,1436), with the same argument lists, so the call sites in
!! prim_driver_mod.F90:686-689,781-784,798-801 and prim_advection_mod.F90:653-656,715-718,1279-1282 stay as they are; only
!! `use cuda_mod` becomes `use tse_mod`.  Build with -DUSE_CUDA_FORTRAN (any Fortran 2003 compiler: the module contains no CUDA
!! Fortran) and link libtse_cuda.so.
!!
!! Every bind(C) interface below is complete (one per entry of include/tse.h, same order); the struct types mirror
!! tse_config / tse_geometry / tse_connectivity / tse_hvcoord field by field.
!!
!! NOT COMPILED IN THE BUILD IMAGE OF THIS REPOSITORY (it has no Fortran compiler).  The C side of the same ABI is compiled and
!! run by tests/c_abi/abi_check.c (C11) and driver/prim_main.cpp; see INTEGRATION.md.
module tse_mod
  use iso_c_binding
  use kinds,          only : real_kind
  use dimensions_mod, only : np, nlev, qsize, qsize_d, nelemd, ne, max_corner_elem
  use element_mod,    only : element_t
  use derivative_mod, only : derivative_t
  use hybvcoord_mod,  only : hvcoord_t
  use hybrid_mod,     only : hybrid_t
  use parallel_mod,   only : abortmp
  implicit none
  private

  ! ---- the reference's hook names (cuda_mod.F90) --------------------------------------------------------------------
  public :: cuda_mod_init, copy_qdp_h2d, copy_qdp_d2h, euler_step_cuda, qdp_time_avg_cuda, vertical_remap_cuda
  public :: advance_hypervis_scalar_cuda
  ! ---- extras of the new library ------------------------------------------------------------------------------------
  public :: tse_shim_finalize, tse_shim_set_derived, tse_shim_precompute_divdp, tse_shim_advec_tracers_remap_rk2
  public :: tse_shim_diag_mass, tse_shim_diag_qminmax, tse_shim_field_hash, tse_shim_synchronize

  type(c_ptr), save :: h = c_null_ptr          !< tse_handle

  ! DSSopt values of prim_advection_mod.F90:454-457 are passed through unchanged (TSE_DSS_* in tse.h)

  type, bind(C) :: tse_config
    integer(c_int) :: ne, nelemd, qsize, qsize_d, nlev, np, rsplit, qsplit
    integer(c_int) :: limiter_option, hypervis_order, hypervis_subcycle_q, vert_remap_q_alg
    real(c_double) :: nu_q
    integer(c_int) :: device
  end type
  type, bind(C) :: tse_geometry
    type(c_ptr) :: spheremp, rspheremp, metdet, rmetdet, Dinv, lat, lon
  end type
  type, bind(C) :: tse_connectivity
    type(c_ptr)    :: putmapP, getmapP, reverse
    integer(c_int) :: nbuf
    type(c_ptr)    :: sfc_index
    integer(c_int) :: ncycles
    type(c_ptr)    :: cyc_rank, cyc_ptr, cyc_len
  end type
  type, bind(C) :: tse_hvcoord
    type(c_ptr)    :: hyai, hybi, hyam, hybm
    real(c_double) :: ps0
  end type

  interface
    function tse_last_error() bind(C, name='tse_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function tse_device_count() bind(C, name='tse_device_count') result(n)
      import :: c_int
      integer(c_int) :: n
    end function
    function tse_init(cfg, geom, conn, hv, dvv, handle) bind(C, name='tse_init') result(rc)
      import :: c_int, c_ptr, c_double, tse_config, tse_geometry, tse_connectivity, tse_hvcoord
      type(tse_config),       intent(in)  :: cfg
      type(tse_geometry),     intent(in)  :: geom
      type(tse_connectivity), intent(in)  :: conn
      type(tse_hvcoord),      intent(in)  :: hv
      real(c_double),         intent(in)  :: dvv(*)
      type(c_ptr),            intent(out) :: handle
      integer(c_int) :: rc
    end function
    function tse_finalize(handle) bind(C, name='tse_finalize') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function tse_synchronize(handle) bind(C, name='tse_synchronize') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function tse_comm_unique_id(id128) bind(C, name='tse_comm_unique_id') result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)
      integer(c_int) :: rc
    end function
    function tse_comm_init(handle, nranks, rank, id128) bind(C, name='tse_comm_init') result(rc)
      import :: c_int, c_ptr, c_char
      type(c_ptr),    value :: handle
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id128(128)
      integer(c_int) :: rc
    end function
    function tse_copy_qdp_h2d(handle, qdp, elem_stride, tl) bind(C, name='tse_copy_qdp_h2d') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, qdp
      integer(c_long_long), value :: elem_stride
      integer(c_int),       value :: tl
      integer(c_int) :: rc
    end function
    function tse_copy_qdp_d2h(handle, qdp, elem_stride, tl) bind(C, name='tse_copy_qdp_d2h') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, qdp
      integer(c_long_long), value :: elem_stride
      integer(c_int),       value :: tl
      integer(c_int) :: rc
    end function
    function tse_set_derived(handle, vn0, s_vn0, dp, s_dp, eta, s_eta, omega, s_omega) bind(C, name='tse_set_derived') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, vn0, dp, eta, omega
      integer(c_long_long), value :: s_vn0, s_dp, s_eta, s_omega
      integer(c_int) :: rc
    end function
    function tse_get_derived(handle, divdp, s_divdp, proj, s_proj, eta, s_eta, omega, s_omega) bind(C, name='tse_get_derived') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, divdp, proj, eta, omega
      integer(c_long_long), value :: s_divdp, s_proj, s_eta, s_omega
      integer(c_int) :: rc
    end function
    function tse_get_dp3d_ps(handle, dp3d, s_dp3d, ps_v, s_ps) bind(C, name='tse_get_dp3d_ps') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, dp3d, ps_v
      integer(c_long_long), value :: s_dp3d, s_ps
      integer(c_int) :: rc
    end function
    function tse_get_qminmax(handle, qmin, qmax) bind(C, name='tse_get_qminmax') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle, qmin, qmax
      integer(c_int) :: rc
    end function
    function tse_precompute_divdp(handle) bind(C, name='tse_precompute_divdp') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function tse_euler_step(handle, np1_qdp, n0_qdp, dt, DSSopt, rhs_multiplier) bind(C, name='tse_euler_step') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      integer(c_int), value :: np1_qdp, n0_qdp, DSSopt, rhs_multiplier
      real(c_double), value :: dt
      integer(c_int) :: rc
    end function
    function tse_qdp_time_avg(handle, rkstage, n0_qdp, np1_qdp) bind(C, name='tse_qdp_time_avg') result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: handle
      integer(c_int), value :: rkstage, n0_qdp, np1_qdp
      integer(c_int) :: rc
    end function
    function tse_vertical_remap(handle, dt, np1, np1_qdp) bind(C, name='tse_vertical_remap') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      real(c_double), value :: dt
      integer(c_int), value :: np1, np1_qdp
      integer(c_int) :: rc
    end function
    function tse_advec_tracers_remap_rk2(handle, dt, nstep) bind(C, name='tse_advec_tracers_remap_rk2') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      real(c_double), value :: dt
      integer(c_int), value :: nstep
      integer(c_int) :: rc
    end function
    function tse_advance_hypervis_scalar(handle, nt_qdp, dt2) bind(C, name='tse_advance_hypervis_scalar') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      integer(c_int), value :: nt_qdp
      real(c_double), value :: dt2
      integer(c_int) :: rc
    end function
    function tse_dcmip_init(handle, test_case) bind(C, name='tse_dcmip_init') result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: handle
      integer(c_int), value :: test_case
      integer(c_int) :: rc
    end function
    function tse_prim_run_subcycle(handle, tstep, nstep) bind(C, name='tse_prim_run_subcycle') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      real(c_double), value :: tstep
      integer(c_int), intent(inout) :: nstep
      integer(c_int) :: rc
    end function
    function tse_diag_mass(handle, tl, mass) bind(C, name='tse_diag_mass') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      integer(c_int), value :: tl
      real(c_double), intent(out) :: mass(*)
      integer(c_int) :: rc
    end function
    function tse_diag_qminmax(handle, tl, qmin, qmax) bind(C, name='tse_diag_qminmax') result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      integer(c_int), value :: tl
      real(c_double), intent(out) :: qmin(*), qmax(*)
      integer(c_int) :: rc
    end function
    function tse_diag_field_hash(handle, tl, hash) bind(C, name='tse_diag_field_hash') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),    value :: handle
      integer(c_int), value :: tl
      integer(c_long_long), intent(out) :: hash(*)      ! unsigned 64-bit patterns
      integer(c_int) :: rc
    end function
    function tse_debug_limiter(n, ptens_w, sphweights, dpmass, minp, maxp) bind(C, name='tse_debug_limiter') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: n
      real(c_double), intent(inout) :: ptens_w(*), minp(*), maxp(*)
      real(c_double), intent(in)    :: sphweights(*), dpmass(*)
      integer(c_int) :: rc
    end function
    function tse_timer_ms(handle, name) bind(C, name='tse_timer_ms') result(ms)
      import :: c_ptr, c_char, c_double
      type(c_ptr), value :: handle
      character(kind=c_char), intent(in) :: name(*)
      real(c_double) :: ms
    end function
    function tse_timer_reset(handle) bind(C, name='tse_timer_reset') result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function tse_mark(handle, slot) bind(C, name='tse_mark') result(rc)
      import :: c_int, c_ptr
      type(c_ptr),    value :: handle
      integer(c_int), value :: slot
      integer(c_int) :: rc
    end function
    function tse_mark_elapsed_ms(handle, a, b) bind(C, name='tse_mark_elapsed_ms') result(ms)
      import :: c_int, c_ptr, c_double
      type(c_ptr),    value :: handle
      integer(c_int), value :: a, b
      real(c_double) :: ms
    end function
    function tse_get_wind(handle, vn0, s_vn0, dp, s_dp) bind(C, name='tse_get_wind') result(rc)
      import :: c_int, c_ptr, c_long_long
      type(c_ptr),          value :: handle, vn0, dp
      integer(c_long_long), value :: s_vn0, s_dp
      integer(c_int) :: rc
    end function
    function tse_stage_launch_count(handle) bind(C, name='tse_stage_launch_count') result(n)
      import :: c_ptr, c_long_long
      type(c_ptr), value :: handle
      integer(c_long_long) :: n
    end function
    function tse_halo_bytes(handle) bind(C, name='tse_halo_bytes') result(n)
      import :: c_ptr, c_long_long
      type(c_ptr), value :: handle
      integer(c_long_long) :: n
    end function
    function tse_launch_count(handle) bind(C, name='tse_launch_count') result(n)
      import :: c_ptr, c_long_long
      type(c_ptr), value :: handle
      integer(c_long_long) :: n
    end function
    function tse_device_bytes(handle) bind(C, name='tse_device_bytes') result(n)
      import :: c_ptr, c_long_long
      type(c_ptr), value :: handle
      integer(c_long_long) :: n
    end function
    function c_strlen(s) bind(C, name='strlen') result(n)
      import :: c_ptr, c_size_t
      type(c_ptr), value :: s
      integer(c_size_t) :: n
    end function
  end interface

contains

  !> abortmp with the library's message (parallel_mod.F90:274-290)
  subroutine check(rc, where)
    integer(c_int),   intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    type(c_ptr) :: p
    integer :: i, n
    if (rc == 0) return
    p = tse_last_error()
    n = min(int(c_strlen(p)), len(msg))
    call c_f_pointer(p, cmsg, [n])
    msg = ' '
    do i = 1, n
      msg(i:i) = cmsg(i)
    end do
    call abortmp(trim(where)//': '//trim(msg))
  end subroutine check

  !> doubles between state%Qdp (or any other component) of consecutive elements of elem(:): the AoS stride of element_t
  function elem_stride(elem) result(s)
    type(element_t), intent(in), target :: elem(:)
    integer(c_long_long) :: s
    if (size(elem) > 1) then
      s = (transfer(c_loc(elem(2)%state%Qdp), 0_c_intptr_t) - transfer(c_loc(elem(1)%state%Qdp), 0_c_intptr_t)) / 8
    else
      s = int(size(elem(1)%state%Qdp), c_long_long)
    end if
  end function elem_stride

  !> cuda_mod_init (cuda_mod.F90:168-410): geometry, connectivity, vertical coordinate -> tse_init; NCCL bootstrap over MPI.
  !! The metric terms are gathered into dense [nelemd] arrays once (they live inside element_t, not contiguous across elements).
  subroutine cuda_mod_init(elem, hybrid, deriv, hvcoord)
    use control_mod,   only : rsplit, qsplit, limiter_option, hypervis_order, hypervis_subcycle_q, vert_remap_q_alg, nu_q
    use schedtype_mod, only : schedule
    use parallel_mod,  only : mpireal_t
#ifdef _MPI
    use mpi
#endif
    type(element_t),    intent(in), target :: elem(:)
    type(hybrid_t),     intent(in) :: hybrid
    type(derivative_t), intent(in), target :: deriv
    type(hvcoord_t),    intent(in), target :: hvcoord
    type(tse_config)       :: cfg
    type(tse_geometry)     :: geom
    type(tse_connectivity) :: conn
    type(tse_hvcoord)      :: hv
    real(c_double), allocatable, target :: spheremp(:,:,:), rspheremp(:,:,:), metdet(:,:,:), rmetdet(:,:,:), Dinv(:,:,:,:,:), lat(:,:,:), lon(:,:,:)
    integer(c_int), allocatable, target :: putmapP(:,:), getmapP(:,:), reverse(:,:), sfc(:), cyc_rank(:), cyc_ptr(:), cyc_len(:)
    character(kind=c_char), target :: id(128)
    integer :: ie, i, j, d, ic, ncyc, ierr

    allocate(spheremp(np,np,nelemd), rspheremp(np,np,nelemd), metdet(np,np,nelemd), rmetdet(np,np,nelemd), Dinv(2,2,np,np,nelemd))
    allocate(lat(np,np,nelemd), lon(np,np,nelemd), putmapP(8,nelemd), getmapP(8,nelemd), reverse(8,nelemd), sfc(nelemd))
    do ie = 1, nelemd
      spheremp(:,:,ie)  = elem(ie)%spheremp
      rspheremp(:,:,ie) = elem(ie)%rspheremp
      metdet(:,:,ie)    = elem(ie)%metdet
      rmetdet(:,:,ie)   = elem(ie)%rmetdet
      Dinv(:,:,:,:,ie)  = elem(ie)%Dinv
      do j = 1, np
        do i = 1, np
          lat(i,j,ie) = elem(ie)%spherep(i,j)%lat
          lon(i,j,ie) = elem(ie)%spherep(i,j)%lon
        end do
      end do
      do d = 1, 8
        putmapP(d,ie) = elem(ie)%desc%putmapP(d)
        getmapP(d,ie) = elem(ie)%desc%getmapP(d)
        reverse(d,ie) = merge(1, 0, elem(ie)%desc%reverse(d))
      end do
      sfc(ie) = elem(ie)%vertex%SpaceCurve
    end do
    ncyc = schedule(1)%ncycles
    allocate(cyc_rank(max(ncyc,1)), cyc_ptr(max(ncyc,1)), cyc_len(max(ncyc,1)))
    do ic = 1, ncyc
      cyc_rank(ic) = schedule(1)%SendCycle(ic)%dest - 1        ! 0-based rank
      cyc_ptr(ic)  = schedule(1)%SendCycle(ic)%ptrP - 1        ! ptrP is the 1-based column of buf (schedule_mod.F90:945-970) -> 0-based offset
      cyc_len(ic)  = schedule(1)%SendCycle(ic)%lengthP
    end do

    cfg%ne = ne;  cfg%nelemd = nelemd;  cfg%qsize = qsize;  cfg%qsize_d = qsize_d;  cfg%nlev = nlev;  cfg%np = np
    cfg%rsplit = rsplit;  cfg%qsplit = qsplit;  cfg%limiter_option = limiter_option;  cfg%hypervis_order = hypervis_order
    cfg%hypervis_subcycle_q = hypervis_subcycle_q;  cfg%vert_remap_q_alg = vert_remap_q_alg;  cfg%nu_q = nu_q;  cfg%device = -1
    geom%spheremp = c_loc(spheremp);  geom%rspheremp = c_loc(rspheremp);  geom%metdet = c_loc(metdet);  geom%rmetdet = c_loc(rmetdet)
    geom%Dinv = c_loc(Dinv);  geom%lat = c_loc(lat);  geom%lon = c_loc(lon)
    conn%putmapP = c_loc(putmapP);  conn%getmapP = c_loc(getmapP);  conn%reverse = c_loc(reverse)
    conn%nbuf = 4*(np+max_corner_elem)*nelemd   ! edge buffer width of initEdgeBuffer (edge_mod.F90:150); the maps index into it
    conn%sfc_index = c_loc(sfc);  conn%ncycles = ncyc
    conn%cyc_rank = c_loc(cyc_rank);  conn%cyc_ptr = c_loc(cyc_ptr);  conn%cyc_len = c_loc(cyc_len)
    hv%hyai = c_loc(hvcoord%hyai);  hv%hybi = c_loc(hvcoord%hybi);  hv%hyam = c_loc(hvcoord%hyam);  hv%hybm = c_loc(hvcoord%hybm)
    hv%ps0 = hvcoord%ps0
    call check(tse_init(cfg, geom, conn, hv, deriv%Dvv, h), 'cuda_mod_init/tse_init')

    if (hybrid%par%nprocs > 1) then     ! ncclUniqueId from rank 0, broadcast over the host's communicator
      if (hybrid%par%rank == 0) call check(tse_comm_unique_id(id), 'tse_comm_unique_id')
#ifdef _MPI
      call MPI_Bcast(id, 128, MPI_BYTE, 0, hybrid%par%comm, ierr)
#endif
      call check(tse_comm_init(h, int(hybrid%par%nprocs, c_int), int(hybrid%par%rank, c_int), id), 'tse_comm_init')
    end if
    ! tse_init copied everything it needs: the gather arrays go out of scope here
  end subroutine cuda_mod_init

  !> copy_qdp_h2d (cuda_mod.F90:429-447)
  subroutine copy_qdp_h2d(elem, nt)
    type(element_t), intent(in), target :: elem(:)
    integer,         intent(in) :: nt
    call check(tse_copy_qdp_h2d(h, c_loc(elem(1)%state%Qdp), elem_stride(elem), int(nt, c_int)), 'copy_qdp_h2d')
  end subroutine copy_qdp_h2d

  !> copy_qdp_d2h (cuda_mod.F90:451-469); also the place where a negative-thickness abort of the remap surfaces
  subroutine copy_qdp_d2h(elem, nt)
    type(element_t), intent(in), target :: elem(:)
    integer,         intent(in) :: nt
    call check(tse_copy_qdp_d2h(h, c_loc(elem(1)%state%Qdp), elem_stride(elem), int(nt, c_int)), 'copy_qdp_d2h')
  end subroutine copy_qdp_d2h

  !> derived%vn0, derived%dp, derived%eta_dot_dpdn, derived%omega_p of this step (the reference's port re-uploads them inside
  !! euler_step_cuda, cuda_mod.F90:526-552); asynchronous, call once per tracer step before the first stage
  subroutine tse_shim_set_derived(elem)
    type(element_t), intent(in), target :: elem(:)
    integer(c_long_long) :: s
    s = elem_stride(elem)
    call check(tse_set_derived(h, c_loc(elem(1)%derived%vn0), s, c_loc(elem(1)%derived%dp), s, &
                               c_loc(elem(1)%derived%eta_dot_dpdn), s, c_loc(elem(1)%derived%omega_p), s), 'tse_set_derived')
  end subroutine tse_shim_set_derived

  !> the divdp loop of Prim_Advec_Tracers_remap_rk2 (prim_advection_mod.F90:614-623)
  subroutine tse_shim_precompute_divdp()
    call check(tse_precompute_divdp(h), 'tse_precompute_divdp')
  end subroutine tse_shim_precompute_divdp

  !> euler_step_cuda (cuda_mod.F90:473-597): same argument list; elem/hvcoord/hybrid/deriv/nets/nete are not needed any more
  !! (state is resident, the library works on all local elements; call from the master thread only, as the reference does).
  subroutine euler_step_cuda(np1_qdp, n0_qdp, dt, elem, hvcoord, hybrid, deriv, nets, nete, DSSopt, rhs_multiplier)
    integer,              intent(in) :: np1_qdp, n0_qdp, nets, nete, DSSopt, rhs_multiplier
    real(kind=real_kind), intent(in) :: dt
    type(element_t),      intent(inout), target :: elem(:)
    type(hvcoord_t),      intent(in) :: hvcoord
    type(hybrid_t),       intent(in) :: hybrid
    type(derivative_t),   intent(in) :: deriv
    integer(c_long_long) :: s
    !$OMP BARRIER
    !$OMP MASTER
    if (rhs_multiplier == 0) then          ! first stage of the step: this step's winds, then divdp (:614-623)
      call tse_shim_set_derived(elem)
      call tse_shim_precompute_divdp()
    end if
    call check(tse_euler_step(h, int(np1_qdp, c_int), int(n0_qdp, c_int), real(dt, c_double), int(DSSopt, c_int), &
                              int(rhs_multiplier, c_int)), 'euler_step_cuda')
    ! the DSS'd level field of this stage back into derived% (the reference's euler_step leaves it there, :943-958)
    s = elem_stride(elem)
    call check(tse_get_derived(h, c_null_ptr, s, c_loc(elem(1)%derived%divdp_proj), s, c_loc(elem(1)%derived%eta_dot_dpdn), s, &
                               c_loc(elem(1)%derived%omega_p), s), 'tse_get_derived')
    !$OMP END MASTER
    !$OMP BARRIER
  end subroutine euler_step_cuda

  !> qdp_time_avg_cuda (cuda_mod.F90:601-622)
  subroutine qdp_time_avg_cuda(elem, rkstage, n0_qdp, np1_qdp, limiter_option, nu_p, nets, nete)
    type(element_t),      intent(inout) :: elem(:)
    real(kind=real_kind), intent(in) :: nu_p
    integer,              intent(in) :: rkstage, n0_qdp, np1_qdp, nets, nete, limiter_option
    !$OMP BARRIER
    !$OMP MASTER
    call check(tse_qdp_time_avg(h, int(rkstage, c_int), int(n0_qdp, c_int), int(np1_qdp, c_int)), 'qdp_time_avg_cuda')
    !$OMP END MASTER
    !$OMP BARRIER
  end subroutine qdp_time_avg_cuda

  !> vertical_remap_cuda (cuda_mod.F90:1436-1560): remaps Qdp(np1_qdp) on the device and returns dp3d(np1), ps_v(np1)
  subroutine vertical_remap_cuda(elem, hvcoord, dt, np1, np1_qdp, nets, nete)
    type(element_t),      intent(inout), target :: elem(:)
    type(hvcoord_t),      intent(in) :: hvcoord
    real(kind=real_kind), intent(in) :: dt
    integer,              intent(in) :: np1, np1_qdp, nets, nete
    integer(c_long_long) :: s
    !$OMP BARRIER
    !$OMP MASTER
    call check(tse_vertical_remap(h, real(dt, c_double), int(np1, c_int), int(np1_qdp, c_int)), 'vertical_remap_cuda')
    s = elem_stride(elem)
    call check(tse_get_dp3d_ps(h, c_loc(elem(1)%state%dp3d(1,1,1,np1)), s, c_loc(elem(1)%state%ps_v(1,1,np1)), s), 'tse_get_dp3d_ps')
    !$OMP END MASTER
    !$OMP BARRIER
  end subroutine vertical_remap_cuda

  !> advance_hypervis_scalar_cuda (cuda_mod.F90:624-718): same argument list; derived%dp / divdp_proj must be on the device
  !! (tse_shim_set_derived + the stages of the step have put them there)
  subroutine advance_hypervis_scalar_cuda(edgeAdv, elem, hvcoord, hybrid, deriv, nt, nt_qdp, nets, nete, dt2)
    use edge_mod, only : EdgeBuffer_t
    type(EdgeBuffer_t),   intent(inout) :: edgeAdv
    type(element_t),      intent(inout), target :: elem(:)
    type(hvcoord_t),      intent(in) :: hvcoord
    type(hybrid_t),       intent(in) :: hybrid
    type(derivative_t),   intent(in) :: deriv
    integer,              intent(in) :: nt, nt_qdp, nets, nete
    real(kind=real_kind), intent(in) :: dt2
    !$OMP BARRIER
    !$OMP MASTER
    call check(tse_advance_hypervis_scalar(h, int(nt_qdp, c_int), real(dt2, c_double)), 'advance_hypervis_scalar_cuda')
    !$OMP END MASTER
    !$OMP BARRIER
  end subroutine advance_hypervis_scalar_cuda

  !> the whole of Prim_Advec_Tracers_remap_rk2 (prim_advection_mod.F90:579-640) in one call; nstep = tl%nstep
  subroutine tse_shim_advec_tracers_remap_rk2(elem, dt, nstep)
    type(element_t),      intent(in), target :: elem(:)
    real(kind=real_kind), intent(in) :: dt
    integer,              intent(in) :: nstep
    call tse_shim_set_derived(elem)
    call check(tse_advec_tracers_remap_rk2(h, real(dt, c_double), int(nstep, c_int)), 'tse_advec_tracers_remap_rk2')
  end subroutine tse_shim_advec_tracers_remap_rk2

  !> global tracer mass (the "Q mass" lines of prim_printstate, prim_state_mod.F90:352-385): already reduced over all ranks
  subroutine tse_shim_diag_mass(tl_qdp, mass)
    integer,              intent(in)  :: tl_qdp
    real(kind=real_kind), intent(out) :: mass(qsize)
    call check(tse_diag_mass(h, int(tl_qdp, c_int), mass), 'tse_diag_mass')
  end subroutine tse_shim_diag_mass

  subroutine tse_shim_diag_qminmax(tl_qdp, qmin, qmax)
    integer,              intent(in)  :: tl_qdp
    real(kind=real_kind), intent(out) :: qmin(qsize), qmax(qsize)
    call check(tse_diag_qminmax(h, int(tl_qdp, c_int), qmin, qmax), 'tse_diag_qminmax')
  end subroutine tse_shim_diag_qminmax

  subroutine tse_shim_field_hash(tl_qdp, hash)
    integer,              intent(in)  :: tl_qdp
    integer(c_long_long), intent(out) :: hash(qsize)
    call check(tse_diag_field_hash(h, int(tl_qdp, c_int), hash), 'tse_diag_field_hash')
  end subroutine tse_shim_field_hash

  subroutine tse_shim_synchronize()
    call check(tse_synchronize(h), 'tse_synchronize')
  end subroutine tse_shim_synchronize

  !> prim_finalize
  subroutine tse_shim_finalize()
    if (c_associated(h)) call check(tse_finalize(h), 'tse_finalize')
    h = c_null_ptr
  end subroutine tse_shim_finalize

end module tse_mod
