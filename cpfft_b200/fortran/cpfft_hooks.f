c     cpfft_hooks.f -- reference-side stubs: the bodies a maintainer of
c     maranGit/CPFFT puts in place of FFT_finite_3d.f:138-147 (model
c     hand-over), drive_eps_sig.f, G_K_dF.f and FFT_nr3.f so that the
c     hot path runs in libcpfft_b200.so through module cpfft_iso_c
c     (cpfft_iso_c.f90).  Fixed form like the reference (it includes
c     the reference's common.main).  Generated from the code blocks of
c     INTEGRATION.md section 4.  Not compilable in this repository's
c     image (no Fortran compiler, no MKL): tests/test_abi.py checks that
c     every cpfft_* symbol used here is bound by the interface module.
c
      subroutine cpfft_model_to_gpu()
      use cpfft_iso_c
c        N, matList, matprp, imatprp, tolNR, tolPCG, maxIter, tstep
      use fft
c        c_array, angle_input, crystal_input
      use crystal_data
      implicit none
      include 'common.main'
      type(cpfft_config) :: cfg
      type(cpfft_material), allocatable :: mats(:)
      type(cpfft_crystal),  allocatable :: crys(:)
      real(8), allocatable :: ang(:,:)
      integer :: m, c, e
      cfg%N = N
      cfg%device = 0
      cfg%rank = 0
      cfg%world = 1
      cfg%maxIter = maxIter
      cfg%tolNR = tolNR
      cfg%tolPCG = tolPCG
      cfg%tstep = tstep
      call cpfft_check( cpfft_create(cfg, cpfft_h), out )
      allocate( mats(nummat), crys(max_crystals), ang(3,N3) )
      do m = 1, nummat
c        1 bilinear, 10 cp, stored as REAL*4 (inmat.f:85)
        mats(m)%type = int(matprp(9,m))
c        REAL*4 slots, passed as REAL*4
        mats(m)%e = matprp(1,m)
        mats(m)%nu = matprp(2,m)
        mats(m)%yld_pt = matprp(5,m)
        mats(m)%tan_e = matprp(4,m)
        mats(m)%beta = matprp(3,m)
c        crystal_type            (inmat.f:236)
        mats(m)%crystal = imatprp(105,m)
c        crystals per material point (inmat.f:201-204)
        mats(m)%n_crystals = imatprp(101,m)
      end do
c        Voce / MTS subset of c_array (mod_crystals.f:142-214)
      do c = 1, max_crystals
        crys(c)%slip_type = c_array(c)%slip_type
        crys(c)%elastic_type = c_array(c)%elastic_type
        crys(c)%h_type = c_array(c)%h_type
        crys(c)%miter = c_array(c)%miter
        crys(c)%alter_mode = merge(1, 0, c_array(c)%alter_mode)
        crys(c)%e = c_array(c)%e
        crys(c)%nu = c_array(c)%nu
        crys(c)%mu = c_array(c)%mu
        crys(c)%harden_n = c_array(c)%harden_n
        crys(c)%theta_0 = c_array(c)%theta_o
        crys(c)%tau_y = c_array(c)%tau_y
        crys(c)%tau_v = c_array(c)%tau_v
        crys(c)%voche_m = c_array(c)%voche_m
        crys(c)%iD_v = c_array(c)%iD_v
        crys(c)%eps_dot_0_y = c_array(c)%eps_dot_o_y
        crys(c)%k_0 = c_array(c)%k_o
        crys(c)%burgers = c_array(c)%b
        crys(c)%atol = c_array(c)%atol
        crys(c)%atol1 = c_array(c)%atol1
        crys(c)%rtol = c_array(c)%rtol
        crys(c)%rtol1 = c_array(c)%rtol1
c        MTS (h_type 2), mod_crystals.f:256-275
        crys(c)%tau_a = c_array(c)%tau_a
        crys(c)%tau_hat_y = c_array(c)%tau_hat_y
        crys(c)%g_0_y = c_array(c)%g_o_y
        crys(c)%tau_hat_v = c_array(c)%tau_hat_v
        crys(c)%g_0_v = c_array(c)%g_o_v
        crys(c)%p_y = c_array(c)%p_y
        crys(c)%q_y = c_array(c)%q_y
        crys(c)%p_v = c_array(c)%p_v
        crys(c)%q_v = c_array(c)%q_v
        crys(c)%boltzman = c_array(c)%boltz
        crys(c)%eps_dot_0_v = c_array(c)%eps_dot_o_v
        crys(c)%mu_0 = c_array(c)%mu_o
        crys(c)%D_0 = c_array(c)%D_o
        crys(c)%T_0 = c_array(c)%t_o
      end do
c        what setup_mm10_rknstr looks up per voxel
      do e = 1, N3
c        (drive_eps_sig.f:571-606)
        ang(1:3,e) = angle_input(e,1,1:3)
      end do
      call cpfft_check( cpfft_set_materials(cpfft_h, nummat, mats,
     &    max_crystals, crys), out )
      call cpfft_check( cpfft_set_voxels(cpfft_h, matList, ang), out )
      end subroutine
c
c     polycrystalline material points (n_crystals > 1 or
c     crystal_input file): in cpfft_model_to_gpu, instead of the
c     single-angle hand-over, pass the tables of read_crystal_data
c     ncmax = size(angle_input, 2)
c     allocate( angm(3,ncmax,N3), cidm(ncmax,N3) )
c     do e = 1, N3
c        setup_mm10_rknstr case 2 (drive_eps_sig.f:744-777)
c       osn = data_offset(e)
c       do c = 1, ncmax
c         angm(1:3,c,e) = angle_input(osn,c,1:3)
c         cidm(c,e) = crystal_input(osn,c)
c       end do
c     end do
c     call cpfft_check( cpfft_set_voxels_taylor(cpfft_h, matList, ncmax,
c    &    angm, cidm), out )
c
      subroutine drive_eps_sig( step, iiter )
      use cpfft_iso_c
      implicit none
      include 'common.main'
      integer :: step, iiter
      call thyme( 2, 1 )
      call cpfft_check( cpfft_drive_eps_sig(cpfft_h, step, iiter), out )
      call thyme( 2, 2 )
      end subroutine
c
c        src/dst: CPFFT_DFM, CPFFT_PN1, CPFFT_B, ...
      subroutine G_K_dF_gpu( src, dst, flgK )
      use cpfft_iso_c
      implicit none
      include 'common.main'
      integer :: src, dst
      logical :: flgK
      call cpfft_check( cpfft_G_K_dF(cpfft_h, src, dst, merge(1, 0,
     &    flgK)), out )
      end subroutine
c
      subroutine FFT_nr3()
      use cpfft_iso_c
      use fft, only: BC_all, isNBC, nstep, out_step
      implicit none
      include 'common.main'
      integer(c_int32_t) :: nbc(9), nr(nstep), cg(64, nstep)
      integer(c_int64_t) :: counters(5)
      real(c_double) :: Pbar(9, nstep), sec(3)
      integer :: s
      nbc = merge(1, 0, isNBC)
c        one call per step keeps ouresult in the loop
      do s = 1, nstep
        call cpfft_check( cpfft_FFT_nr3(cpfft_h, 1, BC_all(1,s), nbc,
     &      nr(s), cg(1,s), 64, Pbar(1,s), sec, counters), out )
c        the reference's own lines, formats 1000-1003
        call print_c_string( cpfft_step_log(cpfft_h), out )
c     (FFT_nr3.f:195-199), composed by the library
        if( out_step(s) ) then
c        Fn1 / urcs_n1 / eps_n1 -> blocks, then ouresult
          call cpfft_download_results()
          call ouresult( s )
        end if
      end do
      end subroutine
c
