! cpfft_iso_c.f90 -- ISO_C_BINDING interface to libcpfft_b200.so (include/cpfft_b200.h).
!
! This is the module a CPFFT maintainer adds to src/ and lists in the makefile before
! FFT_nr3.f / drive_eps_sig.f / G_K_dF.f; INTEGRATION.md shows the replacement bodies of those
! three routines.  It cannot be compiled in this repository's image (no Fortran compiler);
! the same entry points are exercised through ctypes (cpfft_b200/api.py) by the tests.
!
! Every function returns an integer: 0 ok, > 0 a condition on which the reference calls
! die_abort (print + stop, mpi_code.f:15-40), < 0 a CUDA / NCCL / usage error.
module cpfft_iso_c
  use, intrinsic :: iso_c_binding
  implicit none

  ! field ids (cpfft_field)
  integer(c_int), parameter :: CPFFT_FN = 0, CPFFT_FN1 = 1, CPFFT_PN = 2, CPFFT_PN1 = 3,      &
       CPFFT_DFM = 4, CPFFT_B = 5, CPFFT_CG_P = 6, CPFFT_CG_AP = 7, CPFFT_CG_R = 8, CPFFT_K4 = 9, &
       CPFFT_URCS_N = 10, CPFFT_URCS_N1 = 11, CPFFT_EPS_N = 12, CPFFT_EPS_N1 = 13,             &
       CPFFT_ROT_N1 = 14, CPFFT_HIST_N = 15, CPFFT_HIST_N1 = 16, CPFFT_CEP = 17
  ! layouts (cpfft_layout): SOA is the reference's column-major (N3, ncomp)
  integer(c_int), parameter :: CPFFT_LAYOUT_SOA = 0, CPFFT_LAYOUT_AOS = 1

  type, bind(c) :: cpfft_config
     integer(c_int32_t) :: N, device, rank, world, maxIter, pad_
     real(c_double)     :: tolNR, tolPCG, tstep
  end type cpfft_config

  type, bind(c) :: cpfft_material          ! matprp slots, REAL*4 on purpose (mod_fft.f:20)
     integer(c_int32_t) :: type, crystal
     real(c_float)      :: e, nu, beta, tan_e, yld_pt
     integer(c_int32_t) :: n_crystals       ! imatprp(101): crystals per material point (inmat.f:201-204)
  end type cpfft_material

  type, bind(c) :: cpfft_crystal           ! c_array(n), Voce subset (mod_crystals.f:142-214)
     integer(c_int32_t) :: slip_type, elastic_type, h_type, alter_mode, miter, pad_
     real(c_double) :: e, nu, mu, harden_n, theta_0, tau_y, tau_v, voche_m, iD_v,   &
                       eps_dot_0_y, k_0, burgers, atol, atol1, rtol, rtol1
     ! MTS hardening (h_type 2)
     real(c_double) :: tau_a, tau_hat_y, g_0_y, tau_hat_v, g_0_v, p_y, q_y, p_v, q_v,   &
                       boltzman, eps_dot_0_v, mu_0, D_0, T_0
  end type cpfft_crystal

  interface
     integer(c_int) function cpfft_create(cfg, handle) bind(c, name='cpfft_create')
       import :: c_int, c_ptr, cpfft_config
       type(cpfft_config), intent(in) :: cfg
       type(c_ptr), intent(out) :: handle
     end function
     subroutine cpfft_destroy(handle) bind(c, name='cpfft_destroy')
       import :: c_ptr
       type(c_ptr), value :: handle
     end subroutine
     type(c_ptr) function cpfft_last_error(handle) bind(c, name='cpfft_last_error')
       import :: c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_set_materials(handle, nmat, mats, ncry, crys) &
          bind(c, name='cpfft_set_materials')
       import :: c_int, c_ptr, cpfft_material, cpfft_crystal
       type(c_ptr), value :: handle
       integer(c_int), value :: nmat, ncry
       type(cpfft_material), intent(in) :: mats(*)
       type(cpfft_crystal), intent(in) :: crys(*)
     end function
     integer(c_int) function cpfft_set_voxels(handle, matlist, angles_deg) bind(c, name='cpfft_set_voxels')
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: handle
       integer(c_int32_t), intent(in) :: matlist(*)
       real(c_double), intent(in) :: angles_deg(3, *)
     end function
     ! polycrystalline material points (n_crystals > 1): angle_input / crystal_input of
     ! read_crystal_data (mod_crystals.f:2111-2210); Fortran shapes angle_input(3, ncmax, nvox),
     ! crystal_input(ncmax, nvox); pass c_null_ptr-equivalent (an unallocated c_ptr) through
     ! cpfft_set_voxels when every point has one crystal
     integer(c_int) function cpfft_set_voxels_taylor(handle, matlist, ncmax, angles_deg, crystal_ids) &
          bind(c, name='cpfft_set_voxels_taylor')
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: handle
       integer(c_int32_t), intent(in) :: matlist(*)
       integer(c_int), value :: ncmax
       real(c_double), intent(in) :: angles_deg(3, ncmax, *)
       integer(c_int32_t), intent(in) :: crystal_ids(ncmax, *)
     end function
     integer(c_int) function cpfft_set_params(handle, tolNR, tolPCG, maxIter, tstep) &
          bind(c, name='cpfft_set_params')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       real(c_double), value :: tolNR, tolPCG, tstep
       integer(c_int), value :: maxIter
     end function
     integer(c_int) function cpfft_hist_size(handle) bind(c, name='cpfft_hist_size')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int64_t) function cpfft_local_voxels(handle) bind(c, name='cpfft_local_voxels')
       import :: c_int64_t, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_drive_eps_sig(handle, step, iter) bind(c, name='cpfft_drive_eps_sig')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: step, iter
     end function
     integer(c_int) function cpfft_G_K_dF(handle, src, dst, flgK) bind(c, name='cpfft_G_K_dF')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: src, dst, flgK
     end function
     integer(c_int) function cpfft_fftPcg(handle, b, x, tol, iters, relres) bind(c, name='cpfft_fftPcg')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       integer(c_int), value :: b, x
       real(c_double), value :: tol
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: relres
     end function
     integer(c_int) function cpfft_tangent_homo(handle, C_homo) bind(c, name='cpfft_tangent_homo')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       real(c_double), intent(out) :: C_homo(81)
     end function
     integer(c_int) function cpfft_mean_P(handle, Pbar) bind(c, name='cpfft_mean_P')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       real(c_double), intent(out) :: Pbar(9)
     end function
     integer(c_int) function cpfft_update(handle) bind(c, name='cpfft_update')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_FFT_nr3(handle, nstep, BC_all, isNBC, nr_iters, cg_iters, cg_cap, &
          Pbar, seconds, counters) bind(c, name='cpfft_FFT_nr3')
       import :: c_int, c_ptr, c_double, c_int32_t, c_int64_t
       type(c_ptr), value :: handle
       integer(c_int), value :: nstep, cg_cap
       real(c_double), intent(in) :: BC_all(9, *)        ! BC_all(:, step) of mod_fft.f:30
       integer(c_int32_t), intent(in) :: isNBC(9)
       integer(c_int32_t), intent(out) :: nr_iters(*), cg_iters(cg_cap, *)
       real(c_double), intent(out) :: Pbar(9, *), seconds(3)
       integer(c_int64_t), intent(out) :: counters(5)
     end function
     type(c_ptr) function cpfft_step_log(handle) bind(c, name='cpfft_step_log')
       import :: c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_step_counter(handle, next_step, cg_truncated) bind(c, name='cpfft_step_counter')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), intent(out) :: next_step, cg_truncated
     end function
     integer(c_int) function cpfft_field_ncomp(handle, f) bind(c, name='cpfft_field_ncomp')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: f
     end function
     integer(c_int) function cpfft_upload(handle, f, host, layout) bind(c, name='cpfft_upload')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       integer(c_int), value :: f, layout
       real(c_double), intent(in) :: host(*)
     end function
     integer(c_int) function cpfft_download(handle, f, host, layout) bind(c, name='cpfft_download')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       integer(c_int), value :: f, layout
       real(c_double), intent(out) :: host(*)
     end function
     integer(c_int) function cpfft_download_fail_flags(handle, flags) bind(c, name='cpfft_download_fail_flags')
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: handle
       integer(c_int32_t), intent(out) :: flags(*)
     end function
     integer(c_int) function cpfft_download_local_iters(handle, iters2) bind(c, name='cpfft_download_local_iters')
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: handle
       integer(c_int32_t), intent(out) :: iters2(2, *)
     end function
     integer(c_int) function cpfft_material_failures(handle, total, last_sweep) &
          bind(c, name='cpfft_material_failures')
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: handle
       integer(c_int64_t), intent(out) :: total, last_sweep
     end function
     integer(c_int) function cpfft_nccl_unique_id(id128) bind(c, name='cpfft_nccl_unique_id')
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id128(128)
     end function
     integer(c_int) function cpfft_nccl_init(handle, id128) bind(c, name='cpfft_nccl_init')
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: handle
       character(kind=c_char), intent(in) :: id128(128)
     end function
     integer(c_int) function cpfft_fp64_peak(handle, tflops) bind(c, name='cpfft_fp64_peak')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       real(c_double), intent(out) :: tflops
     end function
     integer(c_int) function cpfft_exchange_mode(handle) bind(c, name='cpfft_exchange_mode')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_synchronize(handle) bind(c, name='cpfft_synchronize')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     type(c_ptr) function cpfft_stream(handle) bind(c, name='cpfft_stream')
       import :: c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int64_t) function cpfft_kernel_launches(handle) bind(c, name='cpfft_kernel_launches')
       import :: c_int64_t, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_profile_enable(handle, on) bind(c, name='cpfft_profile_enable')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: on
     end function
     integer(c_int) function cpfft_profile_reset(handle) bind(c, name='cpfft_profile_reset')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     integer(c_int) function cpfft_profile_classes() bind(c, name='cpfft_profile_classes')
       import :: c_int
     end function
     type(c_ptr) function cpfft_profile_name(cls) bind(c, name='cpfft_profile_name')
       import :: c_ptr, c_int
       integer(c_int), value :: cls
     end function
     integer(c_int) function cpfft_profile_get(handle, cls, ms, count) bind(c, name='cpfft_profile_get')
       import :: c_int, c_ptr, c_double, c_int64_t
       type(c_ptr), value :: handle
       integer(c_int), value :: cls
       real(c_double), intent(out) :: ms
       integer(c_int64_t), intent(out) :: count
     end function
  end interface

  type(c_ptr), save :: cpfft_h = c_null_ptr     ! the one analysis of the process (module fft is global too)

contains

  ! print the library's message to unit `iout` and stop: the reference's die_abort convention
  subroutine cpfft_check(rc, iout)
    integer(c_int), intent(in) :: rc
    integer, intent(in) :: iout
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (rc == 0) return
    call c_f_pointer(cpfft_last_error(cpfft_h), msg, [512])
    i = 1
    do while (i <= 512)
       if (msg(i) == c_null_char) exit
       i = i + 1
    end do
    write(iout, '(1x,512a1)') msg(1:i - 1)
    write(iout, '(a,i4)') ' >> cpfft_b200 returned ', rc
    stop
  end subroutine cpfft_check

end module cpfft_iso_c
