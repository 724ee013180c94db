!> ISO_C_BINDING interface to libnekcem_b200.so (include/nekcem_b200.h).
!!
!! This is the thin C-ABI layer BASELINE.json's north_star asks for: NekCEM's host code
!! stays Fortran and calls the hand-written sm_100a kernels through these bindings.  One
!! `bind(C)` interface per entry point of the header; the array ids mirror
!! `enum nekcem_b200_array`.  fortran/cem_maxwell_b200.F shows the three call sites that
!! change in the reference (src/cem_drive.F:161-165, 628; src/cem_maxwell.F:327-345).
!!
!! Build: compile with the reference's own flags (-fdefault-real-8 -fdefault-double-8 or -r8,
!! bin/configurenek:117-123) and link -lnekcem_b200 -lcudart -lnccl (INTEGRATION.md).
!! This image has no Fortran compiler, so this file is shipped unverified by a compiler; the
!! same symbols are exercised through ctypes by tests/test_abi.py.
module nekcem_b200
  use, intrinsic :: iso_c_binding
  implicit none
  private :: c_int, c_double, c_char, c_int64_t, c_int32_t, c_float

  integer(c_int), parameter :: NEKCEM_B200_ABI_VERSION = 1

  ! enum nekcem_b200_array
  integer(c_int), parameter :: NKB_DXM1 = 0, NKB_W3MN = 1,                       &
       NKB_RXMN = 2, NKB_RYMN = 3, NKB_RZMN = 4, NKB_SXMN = 5, NKB_SYMN = 6,      &
       NKB_SZMN = 7, NKB_TXMN = 8, NKB_TYMN = 9, NKB_TZMN = 10, NKB_BMN = 11,     &
       NKB_HBM1 = 12, NKB_EBM1 = 13, NKB_UNXM = 14, NKB_UNYM = 15, NKB_UNZM = 16, &
       NKB_AREAM = 17, NKB_Y_0 = 18, NKB_Y_1 = 19, NKB_Z_0 = 20, NKB_Z_1 = 21,    &
       NKB_HN = 22, NKB_EN = 23, NKB_KHN = 24, NKB_KEN = 25,                      &
       NKB_PERMITTIVITY = 26, NKB_PERMEABILITY = 27, NKB_PMLSIGMA = 28,           &
       NKB_PMLBN = 29, NKB_PMLDN = 30, NKB_KPMLBN = 31, NKB_KPMLDN = 32,           &
       NKB_XMN = 33, NKB_YMN = 34, NKB_ZMN = 35, NKB_YCONDUC = 36

  type, bind(C) :: nekcem_b200_desc
     integer(c_int32_t) :: abi_version, ldim, nx1, nelt, imode, ifupwind, ifpec, ifpml
     integer(c_int32_t) :: device, strict, rank, nranks
  end type nekcem_b200_desc

  interface
     integer(c_int) function nekcem_b200_create(desc, handle) bind(C, name='nekcem_b200_create')
       import :: c_int, nekcem_b200_desc
       type(nekcem_b200_desc), intent(in) :: desc
       integer(c_int), intent(out) :: handle
     end function
     integer(c_int) function nekcem_b200_destroy(handle) bind(C, name='nekcem_b200_destroy')
       import :: c_int
       integer(c_int), value :: handle
     end function
     integer(c_int) function nekcem_b200_set_array(handle, which, host, count) &
          bind(C, name='nekcem_b200_set_array')
       import :: c_int, c_double, c_int64_t
       integer(c_int), value :: handle, which
       real(c_double), intent(in) :: host(*)
       integer(c_int64_t), value :: count
     end function
     integer(c_int) function nekcem_b200_get_array(handle, which, host, count) &
          bind(C, name='nekcem_b200_get_array')
       import :: c_int, c_double, c_int64_t
       integer(c_int), value :: handle, which
       real(c_double), intent(out) :: host(*)
       integer(c_int64_t), value :: count
     end function
     integer(c_int) function nekcem_b200_set_faces(handle, glo_num, nxzfl, cempec, ncempec) &
          bind(C, name='nekcem_b200_set_faces')
       import :: c_int, c_int64_t, c_int32_t
       integer(c_int), value :: handle
       integer(c_int64_t), intent(in) :: glo_num(*)
       integer(c_int64_t), value :: nxzfl
       integer(c_int32_t), intent(in) :: cempec(*)
       integer(c_int32_t), value :: ncempec
     end function
     integer(c_int) function nekcem_b200_set_pml(handle, pmlptr, maxpml) &
          bind(C, name='nekcem_b200_set_pml')
       import :: c_int, c_int32_t
       integer(c_int), value :: handle
       integer(c_int32_t), intent(in) :: pmlptr(*)
       integer(c_int32_t), value :: maxpml
     end function
     integer(c_int) function nekcem_b200_comm_unique_id(id) bind(C, name='nekcem_b200_comm_unique_id')
       import :: c_int, c_char
       character(kind=c_char) :: id(128)
     end function
     integer(c_int) function nekcem_b200_comm_init(handle, id) bind(C, name='nekcem_b200_comm_init')
       import :: c_int, c_char
       integer(c_int), value :: handle
       character(kind=c_char), intent(in) :: id(128)
     end function
     integer(c_int) function nekcem_b200_setup(handle) bind(C, name='nekcem_b200_setup')
       import :: c_int
       integer(c_int), value :: handle
     end function
     integer(c_int) function nekcem_b200_set_incident(handle, ninc, facepts, amp, phase, omega) &
          bind(C, name='nekcem_b200_set_incident')
       import :: c_int, c_int32_t, c_double
       integer(c_int), value :: handle
       integer(c_int32_t), value :: ninc
       integer(c_int32_t), intent(in) :: facepts(*)
       real(c_double), intent(in) :: amp(*), phase(*)
       real(c_double), value :: omega
     end function
     integer(c_int) function nekcem_b200_set_volume_source(handle, comp, profile, amp, omega, phase) &
          bind(C, name='nekcem_b200_set_volume_source')
       import :: c_int, c_double
       integer(c_int), value :: handle, comp
       real(c_double), intent(in) :: profile(*)
       real(c_double), value :: amp, omega, phase
     end function
     integer(c_int) function nekcem_b200_set_option(handle, name, value) &
          bind(C, name='nekcem_b200_set_option')
       import :: c_int, c_char
       integer(c_int), value :: handle, value
       character(kind=c_char), intent(in) :: name(*)
     end function
     integer(c_int) function nekcem_b200_set_time(handle, time, dt) bind(C, name='nekcem_b200_set_time')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), value :: time, dt
     end function
     integer(c_int) function nekcem_b200_get_time(handle, time) bind(C, name='nekcem_b200_get_time')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(out) :: time
     end function
     integer(c_int) function nekcem_b200_step(handle, nsteps) bind(C, name='nekcem_b200_step')
       import :: c_int
       integer(c_int), value :: handle, nsteps
     end function
     integer(c_int) function nekcem_b200_step_streamed(handle, hn_in, en_in, hn_out, en_out) &
          bind(C, name='nekcem_b200_step_streamed')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(in) :: hn_in(*), en_in(*)
       real(c_double), intent(out) :: hn_out(*), en_out(*)
     end function
     integer(c_int) function nekcem_b200_restart_ingest(handle, which, as_double, payload) &
          bind(C, name='nekcem_b200_restart_ingest')
       import :: c_int, c_ptr
       integer(c_int), value :: handle, which, as_double
       type(c_ptr), value :: payload
     end function
     integer(c_int) function nekcem_b200_transport(handle, kind) bind(C, name='nekcem_b200_transport')
       import :: c_int, c_int32_t
       integer(c_int), value :: handle
       integer(c_int32_t), intent(out) :: kind
     end function
     integer(c_int) function nekcem_b200_device_count() bind(C, name='nekcem_b200_device_count')
       import :: c_int
     end function
     integer(c_int) function nekcem_b200_stage(handle, rkstep) bind(C, name='nekcem_b200_stage')
       import :: c_int
       integer(c_int), value :: handle, rkstep
     end function
     integer(c_int) function nekcem_b200_synchronize(handle) bind(C, name='nekcem_b200_synchronize')
       import :: c_int
       integer(c_int), value :: handle
     end function
     integer(c_int) function nekcem_b200_error_sums(handle, exact_hn, exact_en, sumsq, linf) &
          bind(C, name='nekcem_b200_error_sums')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(in) :: exact_hn(*), exact_en(*)
       real(c_double), intent(out) :: sumsq(6), linf(6)
     end function
     integer(c_int) function nekcem_b200_error_sums_mode(handle, kind, k, ph, amp, sumsq, linf) &
          bind(C, name='nekcem_b200_error_sums_mode')
       import :: c_int, c_double, c_int32_t
       integer(c_int), value :: handle
       integer(c_int32_t), intent(in) :: kind(18)
       real(c_double), intent(in) :: k(3), ph(3), amp(6)
       real(c_double), intent(out) :: sumsq(6), linf(6)
     end function
     integer(c_int) function nekcem_b200_last_step_ms(handle, ms, launches) &
          bind(C, name='nekcem_b200_last_step_ms')
       import :: c_int, c_float, c_int64_t
       integer(c_int), value :: handle
       real(c_float), intent(out) :: ms
       integer(c_int64_t), intent(out) :: launches
     end function
     !> Drude / Lorentz ADE state (replaces the per-stage cem_maxwell_drude / cem_maxwell_lorentz
     !> calls of the .usr usersrc, src/cem_maxwell.F:3095-3211)
     integer(c_int) function nekcem_b200_set_drude(handle, jn, kjn, params, dindex, n) &
          bind(C, name='nekcem_b200_set_drude')
       import :: c_int, c_double
       integer(c_int), value :: handle, n
       real(c_double), intent(in) :: jn(*), kjn(*), params(*)
       integer(c_int), intent(in) :: dindex(*)
     end function
     integer(c_int) function nekcem_b200_set_lorentz(handle, jn, kjn, params, lindex, n) &
          bind(C, name='nekcem_b200_set_lorentz')
       import :: c_int, c_double
       integer(c_int), value :: handle, n
       real(c_double), intent(in) :: jn(*), kjn(*), params(*)
       integer(c_int), intent(in) :: lindex(*)
     end function
     integer(c_int) function nekcem_b200_get_ade(handle, jn, kjn) bind(C, name='nekcem_b200_get_ade')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(out) :: jn(*), kjn(*)
     end function
     !> Graphene sheets (replaces the per-stage cem_3d/te/tm_graphene_current calls of the .usr
     !> userfsrc and its srcfh -= fjn(:,:,1), src/cem_maxwell.F:2827-3093).  yconduc may be
     !> c_null_ptr-like absent in C; from Fortran pass COMMON /EMWAVE/ yconduc.
     integer(c_int) function nekcem_b200_set_graphene(handle, fjn, kfjn, params, yconduc, &
          gindex, n) bind(C, name='nekcem_b200_set_graphene')
       import :: c_int, c_double
       integer(c_int), value :: handle, n
       real(c_double), intent(in) :: fjn(*), kfjn(*), params(*), yconduc(*)
       integer(c_int), intent(in) :: gindex(*)
     end function
     integer(c_int) function nekcem_b200_get_graphene(handle, fjn, kfjn) &
          bind(C, name='nekcem_b200_get_graphene')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(inout) :: fjn(*), kfjn(*)
     end function
     !> RK tables of COMMON /RKCOEF/ as rk_storage left them (src/cem_common.F:78-114)
     integer(c_int) function nekcem_b200_set_rk_coefficients(handle, a, b, c) &
          bind(C, name='nekcem_b200_set_rk_coefficients')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(in) :: a(5), b(5), c(6)
     end function
     !> modal filter at the end of every step (q_filter, src/nek5_filter.F); intv from build_new_filter
     integer(c_int) function nekcem_b200_set_filter(handle, intv) &
          bind(C, name='nekcem_b200_set_filter')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(in) :: intv(*)
     end function
     !> big-endian "VECTORS" payload of cem_out for EN (which=0) / HN (1), float32 or float64
     integer(c_int) function nekcem_b200_vtk_payload(handle, which, as_double, out) &
          bind(C, name='nekcem_b200_vtk_payload')
       import :: c_int, c_ptr
       integer(c_int), value :: handle, which, as_double
       type(c_ptr), value :: out
     end function
     !> the two halves of a stage around a caller-provided halo exchange (option external_exchange)
     integer(c_int) function nekcem_b200_stage_pack(handle, rkstep) &
          bind(C, name='nekcem_b200_stage_pack')
       import :: c_int
       integer(c_int), value :: handle, rkstep
     end function
     integer(c_int) function nekcem_b200_stage_compute(handle, rkstep) &
          bind(C, name='nekcem_b200_stage_compute')
       import :: c_int
       integer(c_int), value :: handle, rkstep
     end function
     !> leading dimensions of the reference's storage (SIZE pads lelt): (lpts1,3) fields, a .usr's
     !> (lpts,k) ADE and (lxzfl,3,6) graphene arrays
     integer(c_int) function nekcem_b200_set_leading_dims(handle, lpts, lxzfl) &
          bind(C, name='nekcem_b200_set_leading_dims')
       import :: c_int, c_int64_t
       integer(c_int), value :: handle
       integer(c_int64_t), value :: lpts, lxzfl
     end function
     integer(c_int) function nekcem_b200_set_array_ld(handle, which, host, ld) &
          bind(C, name='nekcem_b200_set_array_ld')
       import :: c_int, c_int64_t, c_double
       integer(c_int), value :: handle, which
       real(c_double), intent(in) :: host(*)
       integer(c_int64_t), value :: ld
     end function
     integer(c_int) function nekcem_b200_get_array_ld(handle, which, host, ld) &
          bind(C, name='nekcem_b200_get_array_ld')
       import :: c_int, c_int64_t, c_double
       integer(c_int), value :: handle, which
       real(c_double), intent(inout) :: host(*)
       integer(c_int64_t), value :: ld
     end function
     !> cem_maxwell_op alone (amult of the exponential / eigenvalue drivers): result in KHN / KEN
     integer(c_int) function nekcem_b200_apply_rhs(handle, rktime) &
          bind(C, name='nekcem_b200_apply_rhs')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), value :: rktime
     end function
     integer(c_int) function nekcem_b200_apply_filter(handle) &
          bind(C, name='nekcem_b200_apply_filter')
       import :: c_int
       integer(c_int), value :: handle
     end function
     integer(c_int) function nekcem_b200_get_rk_coefficients(handle, a, b, c) &
          bind(C, name='nekcem_b200_get_rk_coefficients')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(out) :: a(5), b(5), c(6)
     end function
     !> device slices of peer ipeer (0-based) for an exchange done by the host's CUDA-aware MPI
     integer(c_int) function nekcem_b200_halo_buffers(handle, ipeer, send, recv, count) &
          bind(C, name='nekcem_b200_halo_buffers')
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int), value :: handle
       integer(c_int32_t), value :: ipeer
       type(c_ptr), intent(out) :: send, recv
       integer(c_int64_t), intent(out) :: count
     end function
     integer(c_int) function nekcem_b200_plan_npeers(handle, npeers, nhalo) &
          bind(C, name='nekcem_b200_plan_npeers')
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int), value :: handle
       integer(c_int32_t), intent(out) :: npeers
       integer(c_int64_t), intent(out) :: nhalo
     end function
     integer(c_int) function nekcem_b200_plan_peer(handle, ipeer, peer_rank, count, send_facepts) &
          bind(C, name='nekcem_b200_plan_peer')
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int), value :: handle
       integer(c_int32_t), value :: ipeer
       integer(c_int32_t), intent(out) :: peer_rank
       integer(c_int64_t), intent(out) :: count
       type(c_ptr), value :: send_facepts          ! c_null_ptr: only rank and count
     end function
     !> usersol of the layered-media tests on the device; wave: the 40 reals of
     !> nekcem_b200_planewave in declaration order (see include/nekcem_b200.h)
     integer(c_int) function nekcem_b200_error_sums_planewave(handle, wave, region, inpml, time, &
          sumsq, linf) bind(C, name='nekcem_b200_error_sums_planewave')
       import :: c_int, c_double, c_signed_char
       integer(c_int), value :: handle
       real(c_double), intent(in) :: wave(40)
       integer(c_signed_char), intent(in) :: region(*), inpml(*)
       real(c_double), value :: time
       real(c_double), intent(out) :: sumsq(6), linf(6)
     end function
     integer(c_int) function nekcem_b200_geometry_info(handle, n_const, masses_shared) &
          bind(C, name='nekcem_b200_geometry_info')
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int), value :: handle
       integer(c_int64_t), intent(out) :: n_const
       integer(c_int32_t), intent(out) :: masses_shared
     end function
     integer(c_int) function nekcem_b200_algorithmic_bytes(handle, bytes_per_stage) &
          bind(C, name='nekcem_b200_algorithmic_bytes')
       import :: c_int, c_double
       integer(c_int), value :: handle
       real(c_double), intent(out) :: bytes_per_stage
     end function
  end interface

contains

  !> reference error behaviour: print and exitt (src/nek5_comm_mpi.F:650-692)
  subroutine nkb_check(rc, what)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    interface
       function nekcem_b200_last_error() bind(C, name='nekcem_b200_last_error') result(p)
         import :: c_ptr
         type(c_ptr) :: p
       end function
       function c_strlen(p) bind(C, name='strlen') result(n)
         import :: c_ptr, c_size_t
         type(c_ptr), value :: p
         integer(c_size_t) :: n
       end function
    end interface
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: p
    if (rc /= 0) then
       p = nekcem_b200_last_error()
       call c_f_pointer(p, msg, [c_strlen(p)])
       write (6, *) 'nekcem_b200: ', what, ' failed: ', msg
       call exitt(1)
    end if
  end subroutine nkb_check

end module nekcem_b200
