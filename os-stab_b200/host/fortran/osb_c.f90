!> osb_c -- ISO_C_BINDING interfaces to libosb.so (include/osb.h, OSB_VERSION 2).
!!
!! SOURCE ONLY: no Fortran compiler exists in the build image (SURVEY H2), so this
!! module is syntax-reviewed, not compiled (check it with `gfortran -fsyntax-only
!! -fdefault-integer-8 -fdefault-real-8 osb_c.f90` where a compiler exists); the C++ twin
!! drivers in ../ exercise the same C entry points with the same argument order.
!!
!! The reference is compiled with -fdefault-integer-8 -fdefault-real-8
!! (gcc.mak:78-79): every integer crossing the boundary is therefore declared with
!! an explicit C kind, and default `real`/`complex` (promoted to 8/16 bytes) are
!! passed as real(c_double)/complex(c_double_complex).
!!
!! Conventions of the interface bodies below:
!!   * every body carries its own `implicit none` (interface bodies do not inherit the host's);
!!   * scalars that size an array are declared before the array;
!!   * arguments the C ABI allows to be NULL are `type(c_ptr), value`: pass c_null_ptr to
!!     omit them, c_loc(array) (of a `target` array) to receive them.
module osb_c
  use iso_c_binding
  implicit none

  integer(c_int), parameter :: OSB_FLOW_CONTEBL = 0, OSB_FLOW_CONTE = 1, &
                               OSB_FLOW_ORRSOM = 2, OSB_FLOW_ORRSPACE = 3
  integer(c_int), parameter :: OSB_INT_RK4 = 0, OSB_INT_CK45 = 1
  integer(c_int), parameter :: OSB_DISC_FD = 0, OSB_DISC_CHEB = 1, OSB_DISC_CHEB_NUM = 2
  integer(c_int), parameter :: OSB_ADJ_PENCIL = 0, OSB_ADJ_INVBA = 1
  integer(c_int32_t), parameter :: OSB_ST_CONVERGED = 0, OSB_ST_MAXCOUNT = 1, &
                                   OSB_ST_NONFINITE = 2, OSB_ST_UNVERIFIED = 3

  type, bind(C) :: osb_shoot_opts
    integer(c_int32_t) :: flow, integrator, nstep, maxcount
    real(c_double)     :: testalpha, ymin, ymax, errtol, eigtol
  end type osb_shoot_opts

  interface
    integer(c_int) function osb_version() bind(C, name="osb_version")
      import :: c_int
      implicit none
    end function
    integer(c_int) function osb_create(ctx, device) bind(C, name="osb_create")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device
    end function
    !> one context driving ndev devices from this (single) host thread; every batched call below shards over them
    integer(c_int) function osb_create_multi(ctx, ndev, devices) bind(C, name="osb_create_multi")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: ndev
      type(c_ptr), value :: devices          ! c_null_ptr: devices 0..ndev-1 (ndev=0: all)
    end function
    integer(c_int) function osb_destroy(ctx) bind(C, name="osb_destroy")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx
    end function
    !> C string with the text of the last error of this context (convert with c_f_pointer / a strlen loop)
    type(c_ptr) function osb_last_error(ctx) bind(C, name="osb_last_error")
      import :: c_ptr
      implicit none
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function osb_shoot_defaults(flow, opts) bind(C, name="osb_shoot_defaults")
      import :: c_int, osb_shoot_opts
      implicit none
      integer(c_int), value :: flow
      type(osb_shoot_opts), intent(out) :: opts
    end function
    !> replaces `call SOLVE_BL` (contebl.f:196, orrsom.f:139, orrspace.f:177)
    integer(c_int) function osb_profile_blasius(ctx, variant, integrator, nbl, ymin, ymax, &
                                                f2p, beta, prof) bind(C, name="osb_profile_blasius")
      import :: c_int, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx
      integer(c_int), value :: variant, integrator, nbl
      real(c_double), value :: ymin, ymax, f2p, beta
      type(c_ptr), intent(out) :: prof
    end function
    !> batched SOLVE_BL: nprof (f2p, beta) pairs; tables [nbl+1, nprof] (any of them c_null_ptr)
    integer(c_int) function osb_profile_blasius_batch(ctx, variant, integrator, nbl, ymin, ymax, nprof, f2p, beta, &
                                                      U, Uspl, d2U, d2Uspl, f2p_out, npass) &
        bind(C, name="osb_profile_blasius_batch")
      import :: c_int, c_int64_t, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx
      integer(c_int), value :: variant, integrator, nbl
      real(c_double), value :: ymin, ymax
      integer(c_int64_t), value :: nprof
      real(c_double), intent(in) :: f2p(*), beta(*)
      type(c_ptr), value :: U, Uspl, d2U, d2Uspl   ! real(c_double) (0:nbl, nprof)
      type(c_ptr), value :: f2p_out                ! real(c_double) (nprof)
      type(c_ptr), value :: npass                  ! integer(c_int32_t) (nprof)
    end function
    !> replaces `call SHIFT_BL` (conteucbl.f:236)
    integer(c_int) function osb_profile_shift(ctx, prof, uc, shifted) bind(C, name="osb_profile_shift")
      import :: c_int, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx, prof
      real(c_double), value :: uc
      type(c_ptr), intent(out) :: shifted
    end function
    integer(c_int) function osb_profile_channel(ctx, prof) bind(C, name="osb_profile_channel")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx
      type(c_ptr), intent(out) :: prof
    end function
    !> copies the tables back so the host can write fort.10 exactly as before; every output may be c_null_ptr
    integer(c_int) function osb_profile_get(ctx, prof, ydat, U, Uspl, d2U, d2Uspl, npass, f2p_last) &
        bind(C, name="osb_profile_get")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx, prof
      type(c_ptr), value :: ydat, U, Uspl, d2U, d2Uspl   ! real(c_double) (0:nbl)
      type(c_ptr), value :: npass                        ! integer(c_int32_t)
      type(c_ptr), value :: f2p_last                     ! real(c_double)
    end function
    integer(c_int) function osb_profile_free(ctx, prof) bind(C, name="osb_profile_free")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx, prof
    end function
    !> replaces `call CONTE_IC(...)` (contebl.f:205), `call CONTE(...)` (conte.f:131, orrsom.f:151)
    integer(c_int) function osb_shoot_temporal(ctx, opts, prof, npts, alpha, Re, c, passes, qend, &
                                               status, resid, history, hist_cap) &
        bind(C, name="osb_shoot_temporal")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex, osb_shoot_opts
      implicit none
      type(c_ptr), value :: ctx, prof
      type(osb_shoot_opts), intent(in) :: opts
      integer(c_int64_t), value :: npts
      complex(c_double_complex), intent(in) :: alpha(*)
      real(c_double), intent(in) :: Re(*)
      complex(c_double_complex), intent(inout) :: c(*)
      integer(c_int32_t), intent(out) :: passes(*), qend(*), status(*)
      type(c_ptr), value :: resid          ! c_null_ptr or real(c_double) (2, npts): |err|, |c-cm1|
      type(c_ptr), value :: history        ! c_null_ptr or real(c_double) (4, hist_cap, npts): ctemp, err per pass
      integer(c_int32_t), value :: hist_cap
    end function
    !> (new) the whole (alpha, Re) tensor grid from ONE anchor guess: device-side continuation on a coarse grid,
    !> bilinear seeds, one batched solve.  c, passes, qend, status: (na, nre), alpha index fastest
    integer(c_int) function osb_shoot_grid_temporal(ctx, opts, prof, na, alpha, nre, Re, anchor_alpha, anchor_Re, &
                                                    anchor_c, coarse, c, passes, qend, status, seed, info) &
        bind(C, name="osb_shoot_grid_temporal")
      import :: c_int, c_int32_t, c_ptr, c_double, c_double_complex, osb_shoot_opts
      implicit none
      type(c_ptr), value :: ctx, prof
      type(osb_shoot_opts), intent(in) :: opts
      integer(c_int32_t), value :: na, nre
      real(c_double), intent(in) :: alpha(*), Re(*)
      real(c_double), value :: anchor_alpha, anchor_Re
      complex(c_double_complex), intent(in) :: anchor_c
      integer(c_int32_t), value :: coarse          ! 0: at most 65 coarse nodes per axis
      complex(c_double_complex), intent(out) :: c(*)
      integer(c_int32_t), intent(out) :: passes(*), qend(*), status(*)
      type(c_ptr), value :: seed                   ! c_null_ptr or complex(c_double_complex) (na, nre)
      type(c_ptr), value :: info                   ! c_null_ptr or integer(c_int32_t) (4)
    end function
    !> replaces the omega loop around `call CONTE` (orrspace.f:186-195)
    integer(c_int) function osb_shoot_spatial_lines(ctx, opts, prof, nlines, niter, Re, omega0, &
                                                    domega, alpha0, alpha_out, passes, qend, status, &
                                                    history, hist_cap) &
        bind(C, name="osb_shoot_spatial_lines")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex, osb_shoot_opts
      implicit none
      type(c_ptr), value :: ctx, prof
      type(osb_shoot_opts), intent(in) :: opts
      integer(c_int64_t), value :: nlines
      integer(c_int32_t), value :: niter
      real(c_double), intent(in) :: Re(*), domega(*)
      complex(c_double_complex), intent(in) :: omega0(*), alpha0(*)
      complex(c_double_complex), intent(out) :: alpha_out(niter, *)
      integer(c_int32_t), intent(out) :: passes(niter, *), qend(niter, *), status(niter, *)
      type(c_ptr), value :: history        ! c_null_ptr or real(c_double) (4, hist_cap, niter, nlines) (orrspace.f:485-488)
      integer(c_int32_t), value :: hist_cap
    end function
    !> (new) neutral points Im c(alpha, Re) = 0 by a batched secant iteration on alpha
    integer(c_int) function osb_shoot_neutral(ctx, opts, prof, npts, Re, alpha0, c0, dalpha_rel, tol, maxit, &
                                              alpha_out, c_out, steps, status) &
        bind(C, name="osb_shoot_neutral")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex, osb_shoot_opts
      implicit none
      type(c_ptr), value :: ctx, prof
      type(osb_shoot_opts), intent(in) :: opts
      integer(c_int64_t), value :: npts
      real(c_double), intent(in) :: Re(*), alpha0(*)
      complex(c_double_complex), intent(in) :: c0(*)
      real(c_double), value :: dalpha_rel, tol
      integer(c_int32_t), value :: maxit
      real(c_double), intent(out) :: alpha_out(*)
      complex(c_double_complex), intent(out) :: c_out(*)
      integer(c_int32_t), intent(out) :: steps(*), status(*)
    end function
    !> replaces the "second pass" (shoot.f:762-810): y(i,k) un-normalised + vmax
    integer(c_int) function osb_eigfun(ctx, opts, prof, npts, alpha, Re, eig, omega, y, vmax, qend) &
        bind(C, name="osb_eigfun")
      import :: c_int, c_int64_t, c_ptr, c_double, c_double_complex, osb_shoot_opts
      implicit none
      type(c_ptr), value :: ctx, prof
      type(osb_shoot_opts), intent(in) :: opts
      integer(c_int64_t), value :: npts
      type(c_ptr), value :: alpha          ! temporal flows: complex(c_double_complex) (npts); spatial: c_null_ptr
      real(c_double), intent(in) :: Re(*)
      complex(c_double_complex), intent(in) :: eig(*)
      type(c_ptr), value :: omega          ! spatial flow: complex(c_double_complex) (npts); temporal: c_null_ptr
      complex(c_double_complex), intent(out) :: y(4, *)   ! (4, 0:nstep) per point
      complex(c_double_complex), intent(out) :: vmax(*)
      type(c_ptr), value :: qend           ! c_null_ptr or integer(c_int32_t) (npts)
    end function
    !> replaces MAKE_CHANNEL_METRICS + MAKE_DERIVATIVES + INIT_CHANNEL_PROFILE
    !! (orrfdchan.f:111-113, orrcolchan.f:67-70, orrncbl.f:98-107)
    integer(c_int) function osb_chan_create(ctx, disc, N, chan) bind(C, name="osb_chan_create")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx
      integer(c_int), value :: disc, N          ! disc: 0 FD, 1 Chebyshev, 2 Chebyshev + numerical U''
      type(c_ptr), intent(out) :: chan
    end function
    !> mapped boundary-layer operator of orrcolbl.f / orrfdbl.f (MAKE_DERIVATIVES, MAKE_BL_METRICS, MAKE_MATRIX);
    !> U(0:N) at the grid points replaces the external profile file of INIT_BL_PROFILE
    integer(c_int) function osb_chan_create_bl(ctx, disc, N, lmap, U, merge, chan) bind(C, name="osb_chan_create_bl")
      import :: c_int, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx
      integer(c_int), value :: disc, N, merge
      real(c_double), value :: lmap
      real(c_double), intent(in) :: U(0:*)
      type(c_ptr), intent(out) :: chan
    end function
    integer(c_int) function osb_chan_bl_points(disc, N, lmap, y) bind(C, name="osb_chan_bl_points")
      import :: c_int, c_double
      implicit none
      integer(c_int), value :: disc, N
      real(c_double), value :: lmap
      real(c_double), intent(out) :: y(0:*)
    end function
    integer(c_int) function osb_chan_free(ctx, chan) bind(C, name="osb_chan_free")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx, chan
    end function
    !> eta, u, d2u (0:N); D2, D4 (0:N, 0:N) stored row-major (C order); any of them c_null_ptr
    integer(c_int) function osb_chan_tables(ctx, chan, eta, u, d2u, D2, D4) bind(C, name="osb_chan_tables")
      import :: c_int, c_ptr
      implicit none
      type(c_ptr), value :: ctx, chan, eta, u, d2u, D2, D4
    end function
    !> replaces the alpha loop around MAKE_MATRIX + SOLVE_ORR_SOM (orrfdchan.f:116-121); evec_adj: the left
    !> eigenvector ZGEEV / ZGEGV return with JOBVL = 'V' (columns 4-5 of fort.11), in the form adj_form
    integer(c_int) function osb_chan_eig(ctx, chan, Re, npts, alpha, sigma0, omega, evec, evec_adj, adj_form, &
                                         iters, status) bind(C, name="osb_chan_eig")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex
      implicit none
      type(c_ptr), value :: ctx, chan
      real(c_double), value :: Re
      integer(c_int64_t), value :: npts
      complex(c_double_complex), intent(in) :: alpha(*)
      type(c_ptr), value :: sigma0               ! c_null_ptr: the library picks the mode by the reference's rule
      complex(c_double_complex), intent(out) :: omega(*)
      type(c_ptr), value :: evec                 ! c_null_ptr or c_loc of complex (0:N, npts)
      type(c_ptr), value :: evec_adj             ! c_null_ptr or c_loc of complex (0:N, npts)
      integer(c_int), value :: adj_form          ! OSB_ADJ_PENCIL (orrcolchan) | OSB_ADJ_INVBA (orrfdchan)
      type(c_ptr), value :: iters                ! c_null_ptr or integer(c_int32_t) (npts)
      integer(c_int32_t), intent(out) :: status(*)
    end function
    !> the whole spectrum SOLVE_ORR_SOM sorts before keeping one mode (orrfdchan.f:792-811), ascending in Im(w)
    integer(c_int) function osb_chan_spectrum(ctx, chan, Re, npts, alpha, spectrum, status) &
        bind(C, name="osb_chan_spectrum")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex
      implicit none
      type(c_ptr), value :: ctx, chan
      real(c_double), value :: Re
      integer(c_int64_t), value :: npts
      complex(c_double_complex), intent(in) :: alpha(*)
      complex(c_double_complex), intent(out) :: spectrum(*)
      integer(c_int32_t), intent(out) :: status(*)
    end function
    !> replaces the `do while (abs(IMAG(eigenvalue)).gt.1.0e-8)` Newton loop (orrncbl.f:109-116)
    integer(c_int) function osb_chan_neutral(ctx, chan, nRe, Re, alpha0, sfac, tol, maxit, alpha_out, &
                                             omega, steps, status, trace) bind(C, name="osb_chan_neutral")
      import :: c_int, c_int32_t, c_int64_t, c_ptr, c_double, c_double_complex
      implicit none
      type(c_ptr), value :: ctx, chan
      integer(c_int64_t), value :: nRe
      real(c_double), intent(in) :: Re(*), alpha0(*)
      real(c_double), value :: sfac, tol
      integer(c_int32_t), value :: maxit
      real(c_double), intent(out) :: alpha_out(*)
      complex(c_double_complex), intent(out) :: omega(*)
      integer(c_int32_t), intent(out) :: steps(*), status(*)
      type(c_ptr), value :: trace                ! c_null_ptr or c_loc of real (7, maxit, nRe)
    end function
    !> device FP64 peaks (vector DFMA / tensor DMMA micro-benchmarks), TFLOP/s
    integer(c_int) function osb_measure_fp64_peak(ctx, ms_target, tflops, sm_clock_mhz_est) &
        bind(C, name="osb_measure_fp64_peak")
      import :: c_int, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx
      real(c_double), value :: ms_target
      real(c_double), intent(out) :: tflops
      type(c_ptr), value :: sm_clock_mhz_est     ! c_null_ptr or real(c_double)
    end function
    integer(c_int) function osb_measure_dmma_peak(ctx, ms_target, tflops) bind(C, name="osb_measure_dmma_peak")
      import :: c_int, c_ptr, c_double
      implicit none
      type(c_ptr), value :: ctx
      real(c_double), value :: ms_target
      real(c_double), intent(out) :: tflops
    end function
  end interface
end module osb_c
