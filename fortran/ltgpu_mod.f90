MODULE LTGPU_MOD

!  ISO_C_BINDING interface to libltrans_b200.so, the B200-native replacement of the
!  per-particle time step of LTRANS v.2b (update_particles, LTRANS.f90:707-1614).
!
!  One INTERFACE per entry point of include/ltrans_b200.h (same order), the parameter
!  struct, and three small helpers used by the patched host (see fortran/patches/):
!    ltgpu_fill_params  fills TYPE(ltgpu_params) from PARAM_MOD (the namelists of LTRANS.data)
!    ltgpu_check        prints ltgpu_last_error and STOPs on a non-zero status
!    ltgpu_write_errorlog  drains the per-particle events into ErrorLog.txt (formats 21-29)
!
!  NOT COMPILED in the repository's build image (it has no Fortran compiler); it is the source a
!  maintainer adds to Model/ together with fortran/patches/*.patch.  The same ABI is exercised by
!  tests/abi_harness.c (C) and ltransv.2b_b200/host/binding.py (ctypes).
!
!  Conventions: Fortran arrays are passed as they are stored (column-major, 1-based ids);
!  LOGICAL is never passed, the shims convert to INTEGER(C_INT) with MERGE(1,0,flag).

  USE ISO_C_BINDING
  IMPLICIT NONE
  PUBLIC

  ! status codes
  INTEGER(C_INT), PARAMETER :: LTGPU_OK = 0, LTGPU_E_ARG = 1, LTGPU_E_CUDA = 2,         &
                               LTGPU_E_NODEVICE = 3, LTGPU_E_PARTICLE = 4,              &
                               LTGPU_W_EVENTS_LOST = 5
  INTEGER(C_INT), PARAMETER :: LTGPU_RNG_PHILOX = 1, LTGPU_F32 = 4, LTGPU_F64 = 8

  TYPE, BIND(C) :: ltgpu_params                 ! include/ltrans_b200.h, same order
    INTEGER(C_INT)  :: numpar, dt, idt, us, ws
    REAL(C_FLOAT)   :: hc
    INTEGER(C_INT)  :: Vtransform
    REAL(C_DOUBLE)  :: z0
    INTEGER(C_INT)  :: HTurbOn, VTurbOn
    REAL(C_DOUBLE)  :: ConstantHTurb
    INTEGER(C_INT)  :: Behavior, OpenOceanBoundary, mortality, settlementon
    REAL(C_DOUBLE)  :: deadage, pediage, swimstart, swimslow, swimfast
    REAL(C_DOUBLE)  :: Sgradient, sink, Hswimspeed, Swimdepth
    REAL(C_DOUBLE)  :: twistart, twiend, daylength, Em, Kd, thresh
    INTEGER(C_INT)  :: holesExist, seed
    REAL(C_DOUBLE)  :: PI
    INTEGER(C_INT)  :: ErrorFlag, SaltTempOn, TrackCollisions, FreeSlip
    INTEGER(C_INT)  :: rng_mode, field_dtype, vturb_window_sigs, vturb_fp32_walk
  END TYPE ltgpu_params

  TYPE, BIND(C) :: ltgpu_event
    INTEGER(C_INT) :: particle, code
    REAL(C_DOUBLE) :: time
  END TYPE ltgpu_event

  INTERFACE

    INTEGER(C_INT) FUNCTION ltgpu_create(prm, device, ctx) BIND(C, NAME='ltgpu_create')
      IMPORT
      TYPE(ltgpu_params), INTENT(IN) :: prm
      INTEGER(C_INT), VALUE :: device
      TYPE(C_PTR), INTENT(OUT) :: ctx
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_destroy(ctx) BIND(C, NAME='ltgpu_destroy')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    TYPE(C_PTR) FUNCTION ltgpu_last_error(ctx) BIND(C, NAME='ltgpu_last_error')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_set_grid(ctx, vi,uj,ui,vj, rx,ry,ux,uy,vx,vy, depth,angle,          &
        rho_mask,u_mask,v_mask, SC,CS,SCW,CSW, RE,UE,VE, nRE,nUE,nVE, rAdj,uAdj,vAdj)                 &
        BIND(C, NAME='ltgpu_set_grid')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: vi,uj,ui,vj,nRE,nUE,nVE
      REAL(C_DOUBLE), INTENT(IN) :: rx(*),ry(*),ux(*),uy(*),vx(*),vy(*),depth(*),angle(*),            &
                                    SC(*),CS(*),SCW(*),CSW(*)
      INTEGER(C_INT), INTENT(IN) :: rho_mask(*),u_mask(*),v_mask(*),RE(4,*),UE(4,*),VE(4,*),          &
                                    rAdj(nRE,*),uAdj(nUE,*),vAdj(nVE,*)
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_set_bounds(ctx, nbounds,bnd_x,bnd_y,land, maxbound,bx,by,           &
        maxisland,hx,hy,hid) BIND(C, NAME='ltgpu_set_bounds')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: nbounds,maxbound,maxisland
      REAL(C_DOUBLE), INTENT(IN) :: bnd_x(2,*),bnd_y(2,*),bx(*),by(*),hx(*),hy(*)
      INTEGER(C_INT), INTENT(IN) :: land(*),hid(*)
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_set_habitat(ctx, pedges,polys, hedges,holes,                        &
        npoly,poly_id,poly_start,poly_size,poly_maxdis, nhole,hole_id,hole_start,hole_size,           &
        hole_maxdis, elepoly_ptr,elepoly_idx, polyhole_ptr,polyhole_idx)                              &
        BIND(C, NAME='ltgpu_set_habitat')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: pedges,hedges,npoly,nhole
      REAL(C_DOUBLE), INTENT(IN) :: polys(pedges,*),holes(hedges,*),poly_maxdis(*),hole_maxdis(*)
      INTEGER(C_INT), INTENT(IN) :: poly_id(*),poly_start(*),poly_size(*),hole_id(*),hole_start(*),   &
                                    hole_size(*),elepoly_ptr(*),elepoly_idx(*),polyhole_ptr(*),       &
                                    polyhole_idx(*)
    END FUNCTION

    ! r_ele,u_ele,v_ele: C_LOC of the arrays, or C_NULL_PTR to have the device locate the particles
    INTEGER(C_INT) FUNCTION ltgpu_set_particles(ctx, n, first_id, x,y,z,dob, startpoly,               &
        r_ele,u_ele,v_ele) BIND(C, NAME='ltgpu_set_particles')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: n
      INTEGER(C_INT64_T), VALUE :: first_id
      REAL(C_DOUBLE), INTENT(IN) :: x(*),y(*),z(*),dob(*)
      TYPE(C_PTR), VALUE :: startpoly,r_ele,u_ele,v_ele
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_screen_initial(ctx, counts, bad_particle)                           &
        BIND(C, NAME='ltgpu_screen_initial')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT64_T), INTENT(OUT) :: counts(5), bad_particle
    END FUNCTION

    ! one record in ROMS memory order (node fastest, then level); dtype = LTGPU_F32 / LTGPU_F64;
    ! salt, temp may be C_NULL_PTR
    INTEGER(C_INT) FUNCTION ltgpu_push_hydro(ctx, dtype, zeta,u,v,w,aks,salt,temp)                    &
        BIND(C, NAME='ltgpu_push_hydro')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: dtype
      TYPE(C_PTR), VALUE :: zeta,u,v,w,aks,salt,temp
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_rotate_hydro(ctx) BIND(C, NAME='ltgpu_rotate_hydro')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_step(ctx, p, it) BIND(C, NAME='ltgpu_step')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: p, it
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_run_external(ctx, p) BIND(C, NAME='ltgpu_run_external')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: p
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_sync(ctx, bad_particle) BIND(C, NAME='ltgpu_sync')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), INTENT(OUT) :: bad_particle
    END FUNCTION

    ! every array argument is C_LOC(array) or C_NULL_PTR
    INTEGER(C_INT) FUNCTION ltgpu_fetch(ctx, x,y,z,age,status,salt,temp,hitBottom,hitLand,endpoly,    &
        lifespan,r_ele,u_ele,v_ele) BIND(C, NAME='ltgpu_fetch')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      TYPE(C_PTR), VALUE :: x,y,z,age,status,salt,temp,hitBottom,hitLand,endpoly,lifespan,            &
                            r_ele,u_ele,v_ele
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_fetch_lonlat(ctx, spherical, lonmin, latmin, earth_radius,          &
        lon, lat) BIND(C, NAME='ltgpu_fetch_lonlat')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: spherical
      REAL(C_DOUBLE), VALUE :: lonmin, latmin, earth_radius
      REAL(C_DOUBLE), INTENT(OUT) :: lon(*), lat(*)
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_fetch_sigerr(ctx, count) BIND(C, NAME='ltgpu_fetch_sigerr')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), INTENT(OUT) :: count(*)
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_reset_hits(ctx) BIND(C, NAME='ltgpu_reset_hits')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_stats(ctx, counts) BIND(C, NAME='ltgpu_stats')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT64_T), INTENT(OUT) :: counts(8)
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_drain_events(ctx, buf, cap, n) BIND(C, NAME='ltgpu_drain_events')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      TYPE(ltgpu_event), INTENT(OUT) :: buf(*)
      INTEGER(C_INT), VALUE :: cap
      INTEGER(C_INT), INTENT(OUT) :: n
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_events_lost(ctx, lost) BIND(C, NAME='ltgpu_events_lost')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT64_T), INTENT(OUT) :: lost
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_device_ptr(ctx, which, dptr) BIND(C, NAME='ltgpu_device_ptr')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: which
      TYPE(C_PTR), INTENT(OUT) :: dptr
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_export_device(ctx, which, dst_device)                               &
        BIND(C, NAME='ltgpu_export_device')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: which
      TYPE(C_PTR), VALUE :: dst_device
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_fp64_peak(ctx, tflops) BIND(C, NAME='ltgpu_fp64_peak')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      REAL(C_DOUBLE), INTENT(OUT) :: tflops
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_timer_start(ctx) BIND(C, NAME='ltgpu_timer_start')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_timer_stop(ctx, ms) BIND(C, NAME='ltgpu_timer_stop')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      REAL(C_FLOAT), INTENT(OUT) :: ms
    END FUNCTION

    INTEGER(C_INT) FUNCTION ltgpu_kernel_times(ctx, enable, ms, steps)                                &
        BIND(C, NAME='ltgpu_kernel_times')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
      INTEGER(C_INT), VALUE :: enable
      REAL(C_FLOAT), INTENT(OUT) :: ms(4)
      INTEGER(C_INT64_T), INTENT(OUT) :: steps
    END FUNCTION

    INTEGER(C_INT64_T) FUNCTION ltgpu_launch_count(ctx) BIND(C, NAME='ltgpu_launch_count')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

    TYPE(C_PTR) FUNCTION ltgpu_stream(ctx) BIND(C, NAME='ltgpu_stream')
      IMPORT
      TYPE(C_PTR), VALUE :: ctx
    END FUNCTION

  END INTERFACE

  ! the one context of the run (the reference is a single-process program); multi-GPU runs are one
  ! process per GPU, each with its own slice of the particles (INTEGRATION.md section 4)
  TYPE(C_PTR), SAVE :: gpu_ctx = C_NULL_PTR

CONTAINS

  ! ltgpu_params from the namelists read by getParams (LTRANS.h:45-269)
  SUBROUTINE ltgpu_fill_params(prm)
    USE PARAM_MOD, ONLY: numpar,dt,idt,us,ws,hc,Vtransform,z0,HTurbOn,VTurbOn,ConstantHTurb,   &
      Behavior,OpenOceanBoundary,mortality,settlementon,deadage,pediage,swimstart,swimslow,   &
      swimfast,Sgradient,sink,Hswimspeed,Swimdepth,twistart,twiend,daylength,Em,Kd,thresh,    &
      holesExist,seed,PI,ErrorFlag,SaltTempOn,TrackCollisions,FreeSlip
    TYPE(ltgpu_params), INTENT(OUT) :: prm
    prm%numpar = numpar;  prm%dt = dt;  prm%idt = idt;  prm%us = us;  prm%ws = ws
    prm%hc = hc                                  ! REAL(4) in the reference (LTRANS.h:66)
    prm%Vtransform = Vtransform;  prm%z0 = z0
    prm%HTurbOn = MERGE(1, 0, HTurbOn);  prm%VTurbOn = MERGE(1, 0, VTurbOn)
    prm%ConstantHTurb = ConstantHTurb
    prm%Behavior = Behavior
    prm%OpenOceanBoundary = MERGE(1, 0, OpenOceanBoundary)
    prm%mortality = MERGE(1, 0, mortality);  prm%settlementon = MERGE(1, 0, settlementon)
    prm%deadage = deadage;  prm%pediage = pediage;  prm%swimstart = swimstart
    prm%swimslow = swimslow;  prm%swimfast = swimfast;  prm%Sgradient = Sgradient
    prm%sink = sink;  prm%Hswimspeed = Hswimspeed;  prm%Swimdepth = Swimdepth
    prm%twistart = twistart;  prm%twiend = twiend;  prm%daylength = daylength
    prm%Em = Em;  prm%Kd = Kd;  prm%thresh = thresh
    prm%holesExist = MERGE(1, 0, holesExist);  prm%seed = seed;  prm%PI = PI
    prm%ErrorFlag = ErrorFlag;  prm%SaltTempOn = MERGE(1, 0, SaltTempOn)
    prm%TrackCollisions = MERGE(1, 0, TrackCollisions);  prm%FreeSlip = MERGE(1, 0, FreeSlip)
    prm%rng_mode = LTGPU_RNG_PHILOX      ! counter-based stream keyed on particle id and step
    prm%field_dtype = LTGPU_F32          ! ROMS history files are single precision: lossless
    prm%vturb_window_sigs = 0            ! reference semantics of the VTurb SigErr fall-back
    prm%vturb_fp32_walk = 0              ! the random walk in FP64 like the reference
  END SUBROUTINE ltgpu_fill_params

  ! STOP with the library's message on any status other than OK (there is no CPU fallback)
  SUBROUTINE ltgpu_check(ierr, what)
    INTEGER(C_INT), INTENT(IN) :: ierr
    CHARACTER(LEN=*), INTENT(IN) :: what
    CHARACTER(KIND=C_CHAR), POINTER :: msg(:)
    TYPE(C_PTR) :: cmsg
    INTEGER :: i
    IF (ierr == LTGPU_OK) RETURN
    WRITE(*,*) 'ERROR: ', what, ' returned status ', ierr
    cmsg = ltgpu_last_error(gpu_ctx)
    IF (C_ASSOCIATED(cmsg)) THEN
      CALL C_F_POINTER(cmsg, msg, [512])
      DO i = 1, 512
        IF (msg(i) == C_NULL_CHAR) EXIT
        WRITE(*,'(A)',ADVANCE='NO') msg(i)
      ENDDO
      WRITE(*,*)
    ENDIF
    WRITE(*,*) 'The Program Cannot Continue and Will Terminate'
    STOP
  END SUBROUTINE ltgpu_check

  ! ErrorLog.txt lines of the per-particle events, in the order and formats of the serial loop
  ! (LTRANS.f90:337-354 for the initial checks, :761-775 formats 21-29)
  SUBROUTINE ltgpu_write_errorlog()
    INTEGER, PARAMETER :: cap = 4096
    TYPE(ltgpu_event) :: ev(cap)
    INTEGER(C_INT) :: n, ierr
    INTEGER :: k
    !Error Handling Formats (LTRANS.f90:761-775)
    21 FORMAT ('Particle ',I10,' not in rho element after ',I10,' seconds')
    22 FORMAT ('Particle ',I10,' not in u element after '  ,I10,' seconds')
    23 FORMAT ('Particle ',I10,' not in v element after '  ,I10,' seconds')
    24 FORMAT ('Particle ',I10,' out after 3rd reflection after ',I10,' seconds')
    25 FORMAT ('Particle ',I10,' outside main bounds after intersect_reflect after ',I10,' seconds')
    26 FORMAT ('Particle ',I10,' inside island bounds after intersect_reflect after ',I10,' seconds')
    27 FORMAT ('Particle ',I10,' jumped over rho element after ',I10,' seconds')
    28 FORMAT ('Particle ',I10,' jumped over u element after ',I10,' seconds')
    29 FORMAT ('Particle ',I10,' jumped over v element after ',I10,' seconds')
    DO
      ierr = ltgpu_drain_events(gpu_ctx, ev, cap, n)
      IF (ierr == LTGPU_W_EVENTS_LOST) THEN
        WRITE(*,*) 'WARNING: the device event log overflowed; some ErrorLog.txt lines are missing'
      ELSE
        CALL ltgpu_check(ierr, 'ltgpu_drain_events')
      ENDIF
      IF (n > 0) THEN
        OPEN(210,FILE='ErrorLog.txt',POSITION='APPEND')
        DO k = 1, n
          SELECT CASE (ev(k)%code)
            CASE(11); WRITE(210,"('Particle ',I10,' initially outside main bounds')") ev(k)%particle
            CASE(12); WRITE(210,"('Particle ',I10,' initially inside island bounds')") ev(k)%particle
            CASE(13); WRITE(210,"('Particle ',I10,' initially not in rho element')") ev(k)%particle
            CASE(14); WRITE(210,"('Particle ',I10,' initially not in u element')") ev(k)%particle
            CASE(15); WRITE(210,"('Particle ',I10,' initially not in v element')") ev(k)%particle
            CASE(21); WRITE(210,21) ev(k)%particle, INT(ev(k)%time)
            CASE(22); WRITE(210,22) ev(k)%particle, INT(ev(k)%time)
            CASE(23); WRITE(210,23) ev(k)%particle, INT(ev(k)%time)
            CASE(24); WRITE(210,24) ev(k)%particle, INT(ev(k)%time)
            CASE(25); WRITE(210,25) ev(k)%particle, INT(ev(k)%time)
            CASE(26); WRITE(210,26) ev(k)%particle, INT(ev(k)%time)
            CASE(27); WRITE(210,27) ev(k)%particle, INT(ev(k)%time)
            CASE(28); WRITE(210,28) ev(k)%particle, INT(ev(k)%time)
            CASE(29); WRITE(210,29) ev(k)%particle, INT(ev(k)%time)
          END SELECT
        ENDDO
        CLOSE(210)
      ENDIF
      IF (n < cap) EXIT
    ENDDO
  END SUBROUTINE ltgpu_write_errorlog

END MODULE LTGPU_MOD
