!> Replacement for src/submodules/mlegs_scalar_ops.f90 (and the exchange/init bodies of
!> mlegs_scalar_dist.f90 / mlegs_scalar_init.f90): every `module procedure` declared in
!> src/modules/mlegs_scalar.f90 forwards to the C-ABI entry of the same name.
!> src/modules/mlegs_scalar.f90 itself is kept byte-for-byte.  NOT compiled in the development image.
!>
!> s%e is a Fortran pointer onto library-owned storage (cudaMallocManaged when
!> mlegs_b200_use_managed(1) was called before tfm%init, so the apps' direct reads/writes of s%e --
!> apps/vortical_flow_3d.f90:136-137, 379, 438-443 -- keep working; pages migrate on touch).
submodule (mlegs_scalar) mlegs_scalar_ops_b200
  use, intrinsic :: iso_c_binding
  use mlegs_b200_c
  implicit none

contains

  !> type(scalar) -> C mirror (metadata only; e is the same storage)
  subroutine to_c(s, c)
    class(scalar), intent(in) :: s
    type(c_mlegs_field), intent(out) :: c
    integer :: i
    c%e = c_null_ptr
    if (associated(s%e)) c%e = c_loc(s%e(1,1,1))
    c%glb_sz = s%glb_sz; c%loc_sz = s%loc_sz; c%loc_st = s%loc_st; c%axis_comm = s%axis_comm
    c%ln = s%ln
    c%nrchop_offset = s%nrchop_offset; c%npchop_offset = s%npchop_offset; c%nzchop_offset = s%nzchop_offset
    do i = 1, 3
      c%space(i) = s%space(i:i)
    enddo
    c%space(4) = c_null_char
  end subroutine

  !> C mirror -> type(scalar): exchange/trans may have changed e, loc_sz, loc_st, axis_comm, ln, space
  subroutine from_c(c, s)
    type(c_mlegs_field), intent(in) :: c
    class(scalar), intent(inout) :: s
    integer :: i
    s%glb_sz = c%glb_sz; s%loc_sz = c%loc_sz; s%loc_st = c%loc_st; s%axis_comm = c%axis_comm
    s%ln = c%ln
    s%nrchop_offset = c%nrchop_offset; s%npchop_offset = c%npchop_offset; s%nzchop_offset = c%nzchop_offset
    do i = 1, 3
      s%space(i:i) = c%space(i)
    enddo
    call c_f_pointer(c%e, s%e, s%loc_sz)   ! re-associate like dist:45-47's pointer swap
  end subroutine

  !> scalar_init, mlegs_scalar_init.f90:6-67 (slab layouts only: (1,0,2) physical, (2,1,0) spectral).
  !> glb_sz must be the kit's (nrdim, npdim, nzdim): the library sizes every field from the kit it was initialised
  !> with (tfm%init()), exactly what every app passes here (apps/vortical_flow_3d.f90:81-84).
  module procedure scalar_init
    type(c_mlegs_field) :: c
    if (size(axis_comm) .ne. 3) then
      if (rank_glb .eq. 0) write(*,*) "ERROR: scalar_initialize (B200 path) supports the 3-entry axis_comm form only"
      call MPI_abort(comm_glb, error_flag_comm, MPI_err)
    endif
    if (axis_comm(1) .eq. 1) then
      call b200_check(mlegs_b200_field_alloc(c, 'PPP'//c_null_char))
    else
      call b200_check(mlegs_b200_field_alloc(c, 'FFF'//c_null_char))
    endif
    if (any(c%glb_sz .ne. glb_sz)) then
      if (rank_glb .eq. 0) write(*,*) "ERROR: scalar_initialize: glb_sz differs from the transformation kit's dimensions"
      call MPI_abort(comm_glb, error_flag_comm, MPI_err)
    endif
    call from_c(c, this)
    this%space = 'PPP'                        ! mlegs_scalar_init.f90:65
  end procedure

  !> scalar_chop_offset, mlegs_scalar_init.f90:83-104 (type-bound: the vtable of type(scalar) needs it)
  module procedure scalar_chop_offset
    type(c_mlegs_field) :: c
    integer(c_int) :: iof2_, iof3_
    iof2_ = 0; iof3_ = 0
    if (present(iof2)) iof2_ = int(iof2, c_int)
    if (present(iof3)) iof3_ = int(iof3, c_int)
    call to_c(this, c)
    call b200_check(mlegs_b200_field_chop_offset(c, int(iof1, c_int), iof2_, iof3_))
    this%nrchop_offset = c%nrchop_offset; this%npchop_offset = c%npchop_offset; this%nzchop_offset = c%nzchop_offset
  end procedure

  !> set_comm_grps, dist:371-389 + 508-578.  The device library runs the slab decomposition, i.e. the reference's
  !> process grid with dims = (/ nprocs, 1 /): comm_grps(1) holds every rank, comm_grps(2) is a single-rank group
  !> (exchanges along it are re-labellings).  The host keeps these communicators for its own MPI calls (assemble,
  !> check_stability's allreduce); the data path never uses them.
  module procedure subcomm_cart_2d
    integer :: nproc_comm, key
    if (present(dims)) then
      call MPI_comm_size(comm, nproc_comm, MPI_err)
      if ((dims(1) .ne. nproc_comm .or. dims(2) .ne. 1) .and. (dims(1)*dims(2) .ne. 0)) then
        if (rank_glb .eq. 0) write(*,*) "WARNING: set_comm_grps (B200 path) uses the slab grid (/nprocs, 1/)"
      endif
    endif
    call MPI_comm_dup(comm, comm_glb, MPI_err)
    call MPI_comm_rank(comm_glb, rank_glb, MPI_err)
    call MPI_comm_size(comm_glb, nprocs_glb, MPI_err)
    comm_grps(1) = comm_glb
    rank_grps(1) = rank_glb
    nprocs_grps(1) = nprocs_glb
    key = 0
    call MPI_comm_split(comm_glb, rank_glb, key, comm_grps(2), MPI_err)   ! one rank per colour
    rank_grps(2) = 0
    nprocs_grps(2) = 1
  end procedure

  !> scalar_assemble, dist:70-203: the global array on rank 0 of comm_glb.  Not a hot path: every rank drops its slab
  !> into a zeroed global array and the slabs are summed onto rank 0.
  module procedure scalar_assemble
    type(c_mlegs_field) :: c
    complex(p8), dimension(:,:,:), allocatable, target :: loc
    complex(p8), dimension(:,:,:), allocatable :: part
    integer :: n
    integer(c_int) :: ms
    allocate(array_glb(this%glb_sz(1), this%glb_sz(2), this%glb_sz(3)))
    allocate(part(this%glb_sz(1), this%glb_sz(2), this%glb_sz(3)))
    allocate(loc(this%loc_sz(1), this%loc_sz(2), this%loc_sz(3)))
    call to_c(this, c)
    call b200_check(mlegs_b200_field_download(c, c_loc(loc)))
    call b200_check(mlegs_b200_dist_m_stride(c, ms))   ! azimuthal columns are owned cyclically on several ranks
    part = 0.D0
    part(this%loc_st(1)+1:this%loc_st(1)+this%loc_sz(1), &
         this%loc_st(2)+1:this%loc_st(2)+(this%loc_sz(2)-1)*ms+1:ms, &
         this%loc_st(3)+1:this%loc_st(3)+this%loc_sz(3)) = loc
    n = size(part)
    call MPI_reduce(part, array_glb, n, MPI_double_complex, MPI_sum, 0, comm_glb, MPI_err)
    deallocate(part, loc)
  end procedure

  !> scalar_disassemble, dist:205-368: rank 0's global array handed out slab by slab
  module procedure scalar_disassemble
    type(c_mlegs_field) :: c
    complex(p8), dimension(:,:,:), allocatable, target :: loc
    integer :: n
    integer(c_int) :: ms
    if (.not. allocated(array_glb)) allocate(array_glb(this%glb_sz(1), this%glb_sz(2), this%glb_sz(3)))
    n = size(array_glb)
    call MPI_bcast(array_glb, n, MPI_double_complex, 0, comm_glb, MPI_err)
    allocate(loc(this%loc_sz(1), this%loc_sz(2), this%loc_sz(3)))
    call to_c(this, c)
    call b200_check(mlegs_b200_dist_m_stride(c, ms))
    loc = array_glb(this%loc_st(1)+1:this%loc_st(1)+this%loc_sz(1), &
                    this%loc_st(2)+1:this%loc_st(2)+(this%loc_sz(2)-1)*ms+1:ms, &
                    this%loc_st(3)+1:this%loc_st(3)+this%loc_sz(3))
    call b200_check(mlegs_b200_field_upload(c, c_loc(loc)))
    deallocate(loc)
  end procedure

  !> msave_scalar / mload_scalar, submodules/mlegs_scalar_io.f90:6-250 (defaults: formatted, global)
  module procedure msave_scalar
    type(c_mlegs_field) :: c
    integer(c_int) :: ib, ig
    ib = 0; ig = 1
    if (present(is_binary)) ib = merge(1_c_int, 0_c_int, is_binary)
    if (present(is_global)) ig = merge(1_c_int, 0_c_int, is_global)
    call to_c(s, c)
    call b200_check(mlegs_b200_msave(c, trim(fn)//c_null_char, ib, ig))
  end procedure

  module procedure mload_scalar
    type(c_mlegs_field) :: c
    integer(c_int) :: ib, ig
    ib = 0; ig = 1
    if (present(is_binary)) ib = merge(1_c_int, 0_c_int, is_binary)
    if (present(is_global)) ig = merge(1_c_int, 0_c_int, is_global)
    call to_c(s, c)
    call b200_check(mlegs_b200_mload(trim(fn)//c_null_char, c, ib, ig))
    call from_c(c, s)
  end procedure

  module procedure scalar_dealloc
    type(c_mlegs_field) :: c
    call to_c(this, c)
    call b200_check(mlegs_b200_field_free(c))
    nullify(this%e)
  end procedure

  module procedure scalar_copy                 ! assignment(=), mlegs_scalar_init.f90:106-140
    type(c_mlegs_field) :: cd, cs
    call to_c(this, cd); call to_c(that, cs)
    call b200_check(mlegs_b200_field_copy(cd, cs))
    call from_c(cd, this)
  end procedure

  module procedure scalar_exchange             ! dist:6-67
    type(c_mlegs_field) :: c
    call to_c(this, c)
    call b200_check(mlegs_b200_exchange(c, int(axis_old, c_int), int(axis_new, c_int)))
    call from_c(c, this)
  end procedure

  module procedure trans                       ! ops:157-235
    type(c_mlegs_field) :: c
    call to_c(s, c)
    call b200_check(mlegs_b200_trans(c, space))
    call from_c(c, s)
  end procedure

  module procedure chop                        ! ops:6-41
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_chop(c)); call from_c(c, s)
  end procedure

  module procedure dealias                     ! ops:43-70
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_dealias(c)); call from_c(c, s)
  end procedure

  module procedure svv_filter                  ! ops:72-155
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_svv_filter(c, gain)); call from_c(c, s)
  end procedure

  module procedure calcat0                     ! ops:237-272
    type(c_mlegs_field) :: c
    allocate(calc(s%glb_sz(3)))
    call to_c(s, c); call b200_check(mlegs_b200_calcat0(c, calc))
  end procedure

  module procedure calcat1                     ! ops:274-309
    type(c_mlegs_field) :: c
    allocate(calc(s%glb_sz(3)))
    call to_c(s, c); call b200_check(mlegs_b200_calcat1(c, calc))
  end procedure

  module procedure zeroat1                     ! ops:311-325
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_zeroat1(c)); call from_c(c, s)
  end procedure

  module procedure fftreat                     ! ops:1002-1063
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_fftreat(c)); call from_c(c, s)
  end procedure

  module procedure delsqp                      ! ops:327-366
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_delsqp(c)); call from_c(c, s)
  end procedure

  module procedure idelsqp                     ! ops:368-416
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_idelsqp(c)); call from_c(c, s)
  end procedure

  module procedure xxdx                        ! ops:418-463
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_xxdx(c)); call from_c(c, s)
  end procedure

  module procedure del2h                       ! ops:465-518
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_del2h(c)); call from_c(c, s)
  end procedure

  module procedure del2                        ! ops:520-573
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_del2(c)); call from_c(c, s)
  end procedure

  module procedure idel2_preln                 ! ops:575-669
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_idel2(c, 1_c_int, preln)); call from_c(c, s)
  end procedure

  module procedure idel2_proln                 ! ops:671-760
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_idel2(c, 0_c_int, 0.D0)); call from_c(c, s)
  end procedure

  module procedure ihelm                       ! ops:791-854
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_ihelm(c, alpha)); call from_c(c, s)
  end procedure

  module procedure helm                        ! ops:762-789 (with the write-back the reference omits)
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_helm(c, alpha)); call from_c(c, s)
  end procedure

  module procedure helmp                       ! ops:856-903
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_helmp(c, int(power, c_int), alpha, beta)); call from_c(c, s)
  end procedure

  module procedure ihelmp                      ! ops:905-1000
    type(c_mlegs_field) :: c
    call to_c(s, c); call b200_check(mlegs_b200_ihelmp(c, int(power, c_int), alpha, beta)); call from_c(c, s)
  end procedure

  module procedure fefe                        ! ops:1065-1094
    type(c_mlegs_field) :: c, cn
    call to_c(s, c); call to_c(s_rhs_nonlin, cn)
    call b200_check(mlegs_b200_fefe(c, cn, dt)); call from_c(c, s)
  end procedure

  module procedure abab                        ! ops:1096-1155
    type(c_mlegs_field) :: c, cp, cn, cnp
    integer(c_int) :: flag
    flag = 0
    if (present(is_2nd_svis_p)) flag = merge(1_c_int, 0_c_int, is_2nd_svis_p)
    call to_c(s, c); call to_c(s_p, cp); call to_c(s_rhs_nonlin, cn); call to_c(s_rhs_nonlin_p, cnp)
    call b200_check(mlegs_b200_abab(c, cp, cn, cnp, dt, flag))
    call from_c(c, s); call from_c(cp, s_p); call from_c(cnp, s_rhs_nonlin_p)
  end procedure

  module procedure febe                        ! ops:1157-1198
    type(c_mlegs_field) :: c, cn
    call to_c(s, c); call to_c(s_rhs_nonlin, cn)
    call b200_check(mlegs_b200_febe(c, cn, dt)); call from_c(c, s)
  end procedure

  module procedure abcn                        ! ops:1200-1262
    type(c_mlegs_field) :: c, cp, cn, cnp
    call to_c(s, c); call to_c(s_p, cp); call to_c(s_rhs_nonlin, cn); call to_c(s_rhs_nonlin_p, cnp)
    call b200_check(mlegs_b200_abcn(c, cp, cn, cnp, dt))
    call from_c(c, s); call from_c(cp, s_p); call from_c(cnp, s_rhs_nonlin_p)
  end procedure

  module procedure vector_product              ! ops:1264-1306
    type(c_mlegs_field) :: a(6)
    call to_c(vr, a(1)); call to_c(vp, a(2)); call to_c(vz, a(3))
    call to_c(ur, a(4)); call to_c(up, a(5)); call to_c(uz, a(6))
    call b200_check(mlegs_b200_vecprod(a(1), a(2), a(3), a(4), a(5), a(6)))
  end procedure

  module procedure vector_projection           ! ops:1308-1453
    type(c_mlegs_field) :: a(5)
    call to_c(vr, a(1)); call to_c(vp, a(2)); call to_c(vz, a(3)); call to_c(psi, a(4)); call to_c(chi, a(5))
    call b200_check(mlegs_b200_vec2tp(a(1), a(2), a(3), a(4), a(5)))
    call from_c(a(4), psi); call from_c(a(5), chi)
  end procedure

  module procedure vector_reconstruction       ! ops:1455-1545
    type(c_mlegs_field) :: a(5)
    call to_c(psi, a(1)); call to_c(chi, a(2)); call to_c(vr, a(3)); call to_c(vp, a(4)); call to_c(vz, a(5))
    call b200_check(mlegs_b200_tp2vec(a(1), a(2), a(3), a(4), a(5)))
    call from_c(a(3), vr); call from_c(a(4), vp); call from_c(a(5), vz)
  end procedure

  module procedure curl_vector_reconstruction  ! ops:1547-1560
    type(c_mlegs_field) :: a(5)
    call to_c(psi, a(1)); call to_c(chi, a(2)); call to_c(wr, a(3)); call to_c(wp, a(4)); call to_c(wz, a(5))
    call b200_check(mlegs_b200_tp2curlvec(a(1), a(2), a(3), a(4), a(5)))
    call from_c(a(3), wr); call from_c(a(4), wp); call from_c(a(5), wz)
  end procedure

end submodule
