!> ISO_C_BINDING interface to libmlegs_b200.so (include/mlegs_b200.h).
!> Written against the header; NOT compiled in the development image (no Fortran compiler there).
!> Used by the replacement submodules mlegs_scalar_{init,dist,ops}_b200.f90 (INTEGRATION.md).
module mlegs_b200_c
  use, intrinsic :: iso_c_binding
  implicit none
  public

  !> mlegs_params (include/mlegs_b200.h) <- module globals of modules/mlegs_base.f90
  type, bind(C) :: c_mlegs_params
    integer(c_int) :: nr, np, nz
    integer(c_int) :: nrchop, npchop, nzchop
    real(c_double) :: ell, zlen
    real(c_double) :: visc
    integer(c_int) :: hyperpow
    real(c_double) :: hypervisc
    integer(c_int) :: is_svv
    real(c_double) :: svv_cutoff, svv_target, svv_strength, svv_relax
  end type

  !> mlegs_field <- type(scalar), modules/mlegs_scalar.f90:15-49
  type, bind(C) :: c_mlegs_field
    type(c_ptr)    :: e
    integer(c_int) :: glb_sz(3), loc_sz(3), loc_st(3), axis_comm(3)
    real(c_double) :: ln
    integer(c_int) :: nrchop_offset, npchop_offset, nzchop_offset
    character(kind=c_char) :: space(4)
  end type

  interface
    function mlegs_b200_last_error() bind(C, name='mlegs_b200_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function mlegs_b200_use_managed(on) bind(C, name='mlegs_b200_use_managed') result(rc)
      import :: c_int
      integer(c_int), value :: on
      integer(c_int) :: rc
    end function
    function mlegs_b200_init(p, x, w, lognorm, pf, at0, at1, rank, nranks) &
        bind(C, name='mlegs_b200_init') result(rc)
      import :: c_mlegs_params, c_double, c_int
      type(c_mlegs_params), intent(in) :: p
      real(c_double), intent(in) :: x(*), w(*), lognorm(*), pf(*), at0(*), at1(*)
      integer(c_int), value :: rank, nranks
      integer(c_int) :: rc
    end function
    function mlegs_b200_finalize() bind(C, name='mlegs_b200_finalize') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_alloc(f, space3) bind(C, name='mlegs_b200_field_alloc') result(rc)
      import :: c_mlegs_field, c_char, c_int
      type(c_mlegs_field), intent(inout) :: f
      character(kind=c_char), intent(in) :: space3(*)
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_free(f) bind(C, name='mlegs_b200_field_free') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: f
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_copy(dst, src) bind(C, name='mlegs_b200_field_copy') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: dst
      type(c_mlegs_field), intent(in) :: src
      integer(c_int) :: rc
    end function
    function mlegs_b200_device_sync() bind(C, name='mlegs_b200_device_sync') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    function mlegs_b200_exchange(s, axis_old, axis_new) bind(C, name='mlegs_b200_exchange') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int), value :: axis_old, axis_new
      integer(c_int) :: rc
    end function
    function mlegs_b200_trans(s, to) bind(C, name='mlegs_b200_trans') result(rc)
      import :: c_mlegs_field, c_char, c_int
      type(c_mlegs_field), intent(inout) :: s
      character(kind=c_char), intent(in) :: to(3)
      integer(c_int) :: rc
    end function
    function mlegs_b200_trans_host(host_e, from, to, ln) bind(C, name='mlegs_b200_trans_host') result(rc)
      import :: c_ptr, c_char, c_double, c_int
      type(c_ptr), value :: host_e
      character(kind=c_char), intent(in) :: from(3), to(3)
      real(c_double), value :: ln
      integer(c_int) :: rc
    end function
    !> one-field operators: chop, dealias, zeroat1, delsqp, idelsqp, xxdx, del2h, del2
    function mlegs_b200_chop(s) bind(C, name='mlegs_b200_chop') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    !> fftreat, ops:1002-1063
    function mlegs_b200_fftreat(s) bind(C, name='mlegs_b200_fftreat') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_dealias(s) bind(C, name='mlegs_b200_dealias') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_zeroat1(s) bind(C, name='mlegs_b200_zeroat1') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_delsqp(s) bind(C, name='mlegs_b200_delsqp') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_idelsqp(s) bind(C, name='mlegs_b200_idelsqp') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_xxdx(s) bind(C, name='mlegs_b200_xxdx') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_del2h(s) bind(C, name='mlegs_b200_del2h') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_del2(s) bind(C, name='mlegs_b200_del2') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int) :: rc
    end function
    function mlegs_b200_svv_filter(s, gain) bind(C, name='mlegs_b200_svv_filter') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      real(c_double), intent(inout) :: gain
      integer(c_int) :: rc
    end function
    function mlegs_b200_calcat0(s, out_nz) bind(C, name='mlegs_b200_calcat0') result(rc)
      import :: c_mlegs_field, c_double_complex, c_int
      type(c_mlegs_field), intent(in) :: s
      complex(c_double_complex), intent(out) :: out_nz(*)
      integer(c_int) :: rc
    end function
    function mlegs_b200_calcat1(s, out_nz) bind(C, name='mlegs_b200_calcat1') result(rc)
      import :: c_mlegs_field, c_double_complex, c_int
      type(c_mlegs_field), intent(in) :: s
      complex(c_double_complex), intent(out) :: out_nz(*)
      integer(c_int) :: rc
    end function
    function mlegs_b200_idel2(s, have_preln, preln) bind(C, name='mlegs_b200_idel2') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int), value :: have_preln
      real(c_double), value :: preln
      integer(c_int) :: rc
    end function
    function mlegs_b200_ihelm(s, alpha) bind(C, name='mlegs_b200_ihelm') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      real(c_double), value :: alpha
      integer(c_int) :: rc
    end function
    function mlegs_b200_helmp(s, power, alpha, beta) bind(C, name='mlegs_b200_helmp') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int), value :: power
      real(c_double), value :: alpha, beta
      integer(c_int) :: rc
    end function
    function mlegs_b200_ihelmp(s, power, alpha, beta) bind(C, name='mlegs_b200_ihelmp') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int), value :: power
      real(c_double), value :: alpha, beta
      integer(c_int) :: rc
    end function
    function mlegs_b200_fefe(s, nl, dt) bind(C, name='mlegs_b200_fefe') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      type(c_mlegs_field), intent(in) :: nl
      real(c_double), value :: dt
      integer(c_int) :: rc
    end function
    function mlegs_b200_febe(s, nl, dt) bind(C, name='mlegs_b200_febe') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      type(c_mlegs_field), intent(in) :: nl
      real(c_double), value :: dt
      integer(c_int) :: rc
    end function
    function mlegs_b200_abcn(s, s_p, nl, nl_p, dt) bind(C, name='mlegs_b200_abcn') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s, s_p, nl, nl_p
      real(c_double), value :: dt
      integer(c_int) :: rc
    end function
    function mlegs_b200_vecprod(vr, vp, vz, ur, up, uz) bind(C, name='mlegs_b200_vecprod') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: vr, vp, vz
      type(c_mlegs_field), intent(in) :: ur, up, uz
      integer(c_int) :: rc
    end function
    function mlegs_b200_vec2tp(vr, vp, vz, psi, chi) bind(C, name='mlegs_b200_vec2tp') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: vr, vp, vz
      type(c_mlegs_field), intent(inout) :: psi, chi
      integer(c_int) :: rc
    end function
    function mlegs_b200_tp2vec(psi, chi, vr, vp, vz) bind(C, name='mlegs_b200_tp2vec') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: psi, chi
      type(c_mlegs_field), intent(inout) :: vr, vp, vz
      integer(c_int) :: rc
    end function
    function mlegs_b200_tp2curlvec(psi, chi, wr, wp, wz) bind(C, name='mlegs_b200_tp2curlvec') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: psi, chi
      type(c_mlegs_field), intent(inout) :: wr, wp, wz
      integer(c_int) :: rc
    end function
    !> multi-GPU wiring: the host all-gathers the 64-byte IPC handles (MPI_Allgather) and attaches them
    function mlegs_b200_dist_window(dev_ptr, bytes, handle64) bind(C, name='mlegs_b200_dist_window') result(rc)
      import :: c_ptr, c_size_t, c_signed_char, c_int
      type(c_ptr), intent(out) :: dev_ptr
      integer(c_size_t), intent(out) :: bytes
      integer(c_signed_char), intent(out) :: handle64(64)
      integer(c_int) :: rc
    end function
    function mlegs_b200_dist_attach(handles) bind(C, name='mlegs_b200_dist_attach') result(rc)
      import :: c_signed_char, c_int
      integer(c_signed_char), intent(in) :: handles(*)
      integer(c_int) :: rc
    end function
    function mlegs_b200_dist_detach() bind(C, name='mlegs_b200_dist_detach') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function
    !> MPI_Allreduce(sum) of n host doubles on the library's peer windows (check_stability, vortical_flow_3d.f90:404)
    !> local azimuthal column j of s holds global m = loc_st(2) + stride*(j-1) (cyclic ownership on several ranks)
    function mlegs_b200_dist_m_stride(s, stride) bind(C, name='mlegs_b200_dist_m_stride') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: s
      integer(c_int), intent(out) :: stride
      integer(c_int) :: rc
    end function
    function mlegs_b200_dist_allreduce(buf, n) bind(C, name='mlegs_b200_dist_allreduce') result(rc)
      import :: c_double, c_int
      real(c_double), intent(inout) :: buf(*)
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function
    !> one launch per transform stage (and, on several ranks, one fused exchange) for n scalars of the same state;
    !> `fields` holds c_loc() of the n c_mlegs_field structures
    function mlegs_b200_trans_many(n, fields, to) bind(C, name='mlegs_b200_trans_many') result(rc)
      import :: c_ptr, c_char, c_int
      integer(c_int), value :: n
      type(c_ptr), intent(in) :: fields(*)
      character(kind=c_char), intent(in) :: to(*)
      integer(c_int) :: rc
    end function
    !> n host arrays (ordinary s%e storage), H2D + trans + D2H pipelined over PCIe
    function mlegs_b200_trans_host_batch(n, host_e, from, to, ln) bind(C, name='mlegs_b200_trans_host_batch') result(rc)
      import :: c_ptr, c_char, c_double, c_int
      integer(c_int), value :: n
      type(c_ptr), intent(in) :: host_e(*)
      character(kind=c_char), intent(in) :: from(*), to(*)
      real(c_double), intent(in) :: ln(*)
      integer(c_int) :: rc
    end function
    function mlegs_b200_set_stream(cuda_stream) bind(C, name='mlegs_b200_set_stream') result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: cuda_stream
      integer(c_int) :: rc
    end function
    !> visc, hypervisc, svv_* of the module globals changed (timestep_set / read_input re-run)
    function mlegs_b200_update_params(p) bind(C, name='mlegs_b200_update_params') result(rc)
      import :: c_mlegs_params, c_int
      type(c_mlegs_params), intent(in) :: p
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_zero(f) bind(C, name='mlegs_b200_field_zero') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: f
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_upload(f, host_e) bind(C, name='mlegs_b200_field_upload') result(rc)
      import :: c_mlegs_field, c_ptr, c_int
      type(c_mlegs_field), intent(inout) :: f
      type(c_ptr), value :: host_e
      integer(c_int) :: rc
    end function
    function mlegs_b200_field_download(f, host_e) bind(C, name='mlegs_b200_field_download') result(rc)
      import :: c_mlegs_field, c_ptr, c_int
      type(c_mlegs_field), intent(in) :: f
      type(c_ptr), value :: host_e
      integer(c_int) :: rc
    end function
    !> scalar_chop_offset, mlegs_scalar_init.f90:83-104
    function mlegs_b200_field_chop_offset(f, iof1, iof2, iof3) bind(C, name='mlegs_b200_field_chop_offset') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(inout) :: f
      integer(c_int), value :: iof1, iof2, iof3
      integer(c_int) :: rc
    end function
    !> helm, ops:762-789 (with the write-back the reference omits)
    function mlegs_b200_helm(s, alpha) bind(C, name='mlegs_b200_helm') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      real(c_double), value :: alpha
      integer(c_int) :: rc
    end function
    !> abab, ops:1096-1155
    function mlegs_b200_abab(s, s_p, nl, nl_p, dt, is_2nd_svis_p) bind(C, name='mlegs_b200_abab') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s, s_p, nl, nl_p
      real(c_double), value :: dt
      integer(c_int), value :: is_2nd_svis_p
      integer(c_int) :: rc
    end function
    !> y%e = a*y%e + b*x%e: the whole-array expressions of the time loops (vortical_flow_3d.f90:136-137, 379)
    function mlegs_b200_axpby(y, a, x, b) bind(C, name='mlegs_b200_axpby') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: y
      type(c_mlegs_field), intent(in) :: x
      real(c_double), value :: a, b
      integer(c_int) :: rc
    end function
    !> all(ieee_is_finite(s%e)), check_stability (vortical_flow_3d.f90:397-409)
    function mlegs_b200_is_finite(s, all_finite) bind(C, name='mlegs_b200_is_finite') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: s
      integer(c_int), intent(out) :: all_finite
      integer(c_int) :: rc
    end function
    !> 1: later solves reuse the resident LU factors (default); 0: factor on every call like the reference
    function mlegs_b200_solve_cache(on) bind(C, name='mlegs_b200_solve_cache') result(rc)
      import :: c_int
      integer(c_int), value :: on
      integer(c_int) :: rc
    end function
    !> on-device initial conditions / diagnostics of apps/vortical_flow_3d.f90:258-351, 411-447
    function mlegs_b200_fill_physical(s, re, im) bind(C, name='mlegs_b200_fill_physical') result(rc)
      import :: c_mlegs_field, c_double, c_int
      type(c_mlegs_field), intent(inout) :: s
      real(c_double), value :: re, im
      integer(c_int) :: rc
    end function
    function mlegs_b200_qvort_dist_tp(psi, chi, q, ran_noise, seed) bind(C, name='mlegs_b200_qvort_dist_tp') result(rc)
      import :: c_mlegs_field, c_double, c_long_long, c_int
      type(c_mlegs_field), intent(inout) :: psi, chi
      real(c_double), value :: q, ran_noise
      integer(c_long_long), value :: seed
      integer(c_int) :: rc
    end function
    function mlegs_b200_vort_mag(psi, chi, wr, wp, wz, vormag) bind(C, name='mlegs_b200_vort_mag') result(rc)
      import :: c_mlegs_field, c_int
      type(c_mlegs_field), intent(in) :: psi, chi
      type(c_mlegs_field), intent(inout) :: wr, wp, wz, vormag
      integer(c_int) :: rc
    end function
    !> msave / mload of mlegs_scalar_io.f90 in the reference's binary and formatted layouts (fn is NUL-terminated)
    function mlegs_b200_msave(s, fn, is_binary, is_global) bind(C, name='mlegs_b200_msave') result(rc)
      import :: c_mlegs_field, c_char, c_int
      type(c_mlegs_field), intent(in) :: s
      character(kind=c_char), intent(in) :: fn(*)
      integer(c_int), value :: is_binary, is_global
      integer(c_int) :: rc
    end function
    function mlegs_b200_mload(fn, s, is_binary, is_global) bind(C, name='mlegs_b200_mload') result(rc)
      import :: c_mlegs_field, c_char, c_int
      character(kind=c_char), intent(in) :: fn(*)
      type(c_mlegs_field), intent(inout) :: s
      integer(c_int), value :: is_binary, is_global
      integer(c_int) :: rc
    end function
  end interface

contains

  !> turn a non-zero return code into the reference's own `stop '<message>'`
  subroutine b200_check(rc)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    integer :: i
    if (rc .eq. 0) return
    call c_f_pointer(mlegs_b200_last_error(), cmsg, [512])
    msg = ' '
    do i = 1, 512
      if (cmsg(i) .eq. c_null_char) exit
      msg(i:i) = cmsg(i)
    enddo
    write(*,*) trim(msg)
    error stop 1
  end subroutine

end module mlegs_b200_c
