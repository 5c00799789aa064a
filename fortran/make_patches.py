#!/usr/bin/env python
"""Generates fortran/patches/*.patch: the changes a maintainer applies to the UNMODIFIED
LTRANS v.2b sources (Model/) to run the particle loop on libltrans_b200.so.

    python fortran/make_patches.py /path/to/LTRANSv.2b      # writes fortran/patches/
    cd /path/to/LTRANSv.2b && for p in .../fortran/patches/*.patch; do patch -p1 < $p; done
    cp .../fortran/ltgpu_mod.f90 Model/ && cd Model && make LTRANS_gpu LTGPU_LIBDIR=.../ltransv.2b_b200/csrc

Every edit is an exact-text replacement that must match once, so the script fails loudly if the
reference differs from v.2b.  The patches contain only the changed hunks with their context lines.
"""
import difflib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def edit(src, old, new, count=1):
    assert src.count(old) == count, (src.count(old), old[:70])
    return src.replace(old, new)


# ------------------------------------------------------------------ hydrodynamic_module.f90
def hydro(s):
    s = edit(s, """  INTEGER, ALLOCATABLE, DIMENSION(:) :: P_r_element,P_u_element,P_v_element
""", """  INTEGER, ALLOCATABLE, DIMENSION(:) :: P_r_element,P_u_element,P_v_element
  !(the B200 offload locates the particles itself; these stay for setEle_all's start-up check)
""")
    s = edit(s, """    getR_ele,getP_r_element,finHydro,initNetCDF,createNetCDF,writeNetCDF
""", """    getR_ele,getP_r_element,finHydro,initNetCDF,createNetCDF,writeNetCDF,      &
    gpuUploadGrid,gpuSetParticles,gpuPushHydro
""")
    s = edit(s, """    ALLOCATE(v_Adjacent(v_elements,10))
""", """    ALLOCATE(v_Adjacent(v_elements,10))
    !rows are only partly filled below and a 0 entry means "no more neighbours" (setEle):
    !  the offload needs them zero-filled, not uninitialised
    r_Adjacent = 0
    u_Adjacent = 0
    v_Adjacent = 0
""")
    s = edit(s, """END MODULE HYDRO_MOD""", """
  !*********************************************************
  !*              B200 offload (LTGPU_MOD)                 *
  !*********************************************************

  SUBROUTINE gpuUploadGrid()
    !Hands the grid tables built by initGrid to the device (ltgpu_set_grid)
    USE PARAM_MOD, ONLY: ui,vi,uj,vj,rho_elements,u_elements,v_elements
    USE LTGPU_MOD
    IMPLICIT NONE
    CALL ltgpu_check( ltgpu_set_grid(gpu_ctx, vi,uj,ui,vj, rx,ry,ux,uy,vx,vy, depth,rho_angle, &
         rho_mask,u_mask,v_mask, SC,CS,SCW,CSW, RE,UE,VE,                                      &
         rho_elements,u_elements,v_elements, r_Adjacent,u_Adjacent,v_Adjacent),                &
         'ltgpu_set_grid' )
  END SUBROUTINE gpuUploadGrid

  SUBROUTINE gpuSetParticles(x,y,z,dob)
    !Particle table -> device.  The element search of setEle_all is repeated on the device
    !  (C_NULL_PTR for the three element arrays): same result, through a bucket index.
    USE PARAM_MOD, ONLY: numpar
    USE LTGPU_MOD
    IMPLICIT NONE
    DOUBLE PRECISION, INTENT(IN) :: x(numpar),y(numpar),z(numpar),dob(numpar)
    CALL ltgpu_check( ltgpu_set_particles(gpu_ctx, numpar, 1_C_INT64_T, x,y,z,dob,             &
         C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR), 'ltgpu_set_particles' )
  END SUBROUTINE gpuSetParticles

  SUBROUTINE gpuPushHydro(which)
    !Queues one hydrodynamic record for the device: which = 1 back, 2 centre, 3 forward slot of
    !  the t_* arrays (after initHydro: all three; after updateHydro: 3).  The slices are in ROMS
    !  memory order (node fastest, then level) and already multiplied by the masks; the library
    !  copies them into its pinned staging buffer before it returns.
    USE PARAM_MOD, ONLY: us,ws,rho_nodes,u_nodes,v_nodes
    USE LTGPU_MOD
    IMPLICIT NONE
    INTEGER, INTENT(IN) :: which
    INTEGER :: s
    DOUBLE PRECISION, ALLOCATABLE, TARGET :: gz(:),gu(:,:),gv(:,:),gw(:,:),gk(:,:),gs(:,:),gt(:,:)
    s = t_b
    IF (which == 2) s = t_c
    IF (which == 3) s = t_f
    ALLOCATE(gz(rho_nodes),gu(u_nodes,us),gv(v_nodes,us),gw(rho_nodes,ws),gk(rho_nodes,ws))
    ALLOCATE(gs(rho_nodes,us),gt(rho_nodes,us))
    gz = t_zeta(s,:)
    gu = t_Uvel(s,:,:)
    gv = t_Vvel(s,:,:)
    gw = t_Wvel(s,:,:)
    gk = t_Kh(s,:,:)
    gs = t_salt(s,:,:)
    gt = t_temp(s,:,:)
    CALL ltgpu_check( ltgpu_push_hydro(gpu_ctx, LTGPU_F64, C_LOC(gz),C_LOC(gu),C_LOC(gv),      &
         C_LOC(gw),C_LOC(gk),C_LOC(gs),C_LOC(gt)), 'ltgpu_push_hydro' )
    DEALLOCATE(gz,gu,gv,gw,gk,gs,gt)
  END SUBROUTINE gpuPushHydro

END MODULE HYDRO_MOD""")
    return s


# --------------------------------------------------------------------- boundary_module.f90
def boundary(s):
    s = edit(s, """  PUBLIC :: isBndSet,createBounds,mbounds,ibounds,intersect_reflect
""", """  PUBLIC :: isBndSet,createBounds,mbounds,ibounds,intersect_reflect,gpuUploadBounds
""")
    i = s.rindex("END MODULE")
    s = s[:i] + """
  SUBROUTINE gpuUploadBounds()
    !Hands the boundary tables built by createBounds to the device (ltgpu_set_bounds);
    !  LOGICAL land is converted to INTEGER (LOGICAL is not interoperable)
    USE LTGPU_MOD
    IMPLICIT NONE
    INTEGER(C_INT), ALLOCATABLE :: iland(:),ihid(:)
    DOUBLE PRECISION, ALLOCATABLE :: ghx(:),ghy(:)
    INTEGER :: nh
    ALLOCATE(iland(nbounds))
    iland = MERGE(1, 0, land)
    nh = MAX(maxisland,1)
    ALLOCATE(ghx(nh),ghy(nh),ihid(nh))
    ghx = 0.0
    ghy = 0.0
    ihid = 0
    IF (maxisland > 0) THEN
      ghx = hx(1:maxisland)
      ghy = hy(1:maxisland)
      ihid = hid(1:maxisland)
    ENDIF
    CALL ltgpu_check( ltgpu_set_bounds(gpu_ctx, nbounds,bnd_x,bnd_y,iland, maxbound,bx,by,     &
         maxisland,ghx,ghy,ihid), 'ltgpu_set_bounds' )
    DEALLOCATE(iland,ghx,ghy,ihid)
  END SUBROUTINE gpuUploadBounds

""" + s[i:]
    return s


# -------------------------------------------------------------------- settlement_module.f90
def settlement(s):
    s = edit(s, """  PUBLIC :: initSettlement,testSettlement,isSettled,finSettlement
""", """  PUBLIC :: initSettlement,testSettlement,isSettled,finSettlement,gpuUploadHabitat
""")
    s = edit(s, """END MODULE SETTLEMENT_MOD""", """
  SUBROUTINE gpuUploadHabitat()
    !Hands the habitat tables of getHabitat / createPolySpecs to the device (ltgpu_set_habitat):
    !  polys, holes as they are; the id-indexed specs as dense lists; the ragged lists
    !  elepolys(:)%poly, polyholes(:)%poly flattened to CSR with 0-based list positions
    USE PARAM_MOD, ONLY: rho_elements,minpolyid,maxpolyid,minholeid,maxholeid,pedges,hedges,   &
                         holesExist
    USE LTGPU_MOD
    IMPLICIT NONE
    INTEGER :: i,j,k,e,np,nh,ne,nq
    INTEGER(C_INT), ALLOCATABLE :: pid(:),pstart(:),psize(:),hidl(:),hstart(:),hsize(:),       &
                                   ppos(:),hpos(:),eptr(:),eidx(:),qptr(:),qidx(:)
    DOUBLE PRECISION, ALLOCATABLE :: pmax(:),hmax(:),gholes(:,:)

    np = COUNT(polyspecs(:,1) /= 0)
    ALLOCATE(pid(np),pstart(np),psize(np),pmax(np),ppos(minpolyid:maxpolyid))
    ppos = -1
    k = 0
    DO i = minpolyid,maxpolyid
      IF (polyspecs(i,1) == 0) CYCLE
      k = k + 1
      pid(k) = i
      pstart(k) = polyspecs(i,1)
      psize(k) = polyspecs(i,2)
      pmax(k) = maxbdis(i)
      ppos(i) = k - 1
    ENDDO

    nh = 0
    IF (holesExist) nh = COUNT(holespecs(:,1) /= 0)
    ALLOCATE(hidl(MAX(nh,1)),hstart(MAX(nh,1)),hsize(MAX(nh,1)),hmax(MAX(nh,1)))
    ALLOCATE(hpos(minholeid:maxholeid))
    hpos = -1
    IF (nh > 0) THEN
      k = 0
      DO i = minholeid,maxholeid
        IF (holespecs(i,1) == 0) CYCLE
        k = k + 1
        hidl(k) = i
        hstart(k) = holespecs(i,1)
        hsize(k) = holespecs(i,2)
        hmax(k) = maxhdis(i)
        hpos(i) = k - 1
      ENDDO
    ENDIF

    !polygons per rho element
    ne = 0
    DO e = 1,rho_elements
      ne = ne + elepolys(e)%numpoly
    ENDDO
    ALLOCATE(eptr(rho_elements+1),eidx(MAX(ne,1)))
    eptr(1) = 0
    k = 0
    DO e = 1,rho_elements
      DO j = 1,elepolys(e)%numpoly
        k = k + 1
        eidx(k) = ppos(elepolys(e)%poly(j))
      ENDDO
      eptr(e+1) = k
    ENDDO

    !holes per polygon, in the order of the dense polygon list
    nq = 0
    IF (nh > 0) THEN
      DO i = 1,np
        nq = nq + polyholes(pid(i))%numpoly
      ENDDO
    ENDIF
    ALLOCATE(qptr(np+1),qidx(MAX(nq,1)))
    qptr(1) = 0
    k = 0
    DO i = 1,np
      IF (nh > 0) THEN
        DO j = 1,polyholes(pid(i))%numpoly
          k = k + 1
          qidx(k) = hpos(polyholes(pid(i))%poly(j))
        ENDDO
      ENDIF
      qptr(i+1) = k
    ENDDO

    ALLOCATE(gholes(MAX(hedges,1),6))
    gholes = 0.0
    IF (nh > 0) gholes(1:hedges,:) = holes(1:hedges,:)
    CALL ltgpu_check( ltgpu_set_habitat(gpu_ctx, pedges,polys, MERGE(hedges,1,nh>0),gholes,    &
         np,pid,pstart,psize,pmax, nh,hidl,hstart,hsize,hmax, eptr,eidx, qptr,qidx),           &
         'ltgpu_set_habitat' )
    DEALLOCATE(pid,pstart,psize,pmax,ppos,hidl,hstart,hsize,hmax,hpos,eptr,eidx,qptr,qidx,gholes)
  END SUBROUTINE gpuUploadHabitat

END MODULE SETTLEMENT_MOD""")
    return s


# ------------------------------------------------------------------------------ LTRANS.f90
def ltrans(s):
    # state kept by the host for the writers
    s = edit(s, """  INTEGER, ALLOCATABLE, DIMENSION(:) :: startpoly,endpoly,hitBottom,hitLand
""", """  INTEGER, ALLOCATABLE, DIMENSION(:) :: startpoly,endpoly,hitBottom,hitLand
  !B200 offload: getStatus() of every particle as fetched from the device (the behaviour
  !  state lives there), filled by gpu_fetch() before every write
  INTEGER, ALLOCATABLE, TARGET, DIMENSION(:) :: gpu_status
""")
    # ini_LTRANS: create the context and upload after the three initial records are read
    s = edit(s, """    !Read in initial hydrodynamic model data
    CALL initHydro()
""", """    !Read in initial hydrodynamic model data
    CALL initHydro()

    !B200 offload: device context, tables, particles, the three initial records
    CALL gpu_init()
""")
    # run_External_Timestep: push + rotate after updateHydro; no CPU particle loop
    s = edit(s, """      IF(p > 2) CALL updateHydro()   !do not start updating until 3rd iteration
""", """      IF(p > 2) THEN                 !do not start updating until 3rd iteration
        CALL updateHydro()
        CALL gpuPushHydro(3)         !the record just read becomes "forward" on the device
        CALL ltgpu_check( ltgpu_rotate_hydro(gpu_ctx), 'ltgpu_rotate_hydro' )
      ENDIF
""")
    s = edit(s, """    use hydro_mod, only: updateHydro
    integer :: stepIT
""", """    use hydro_mod, only: updateHydro,gpuPushHydro
    use ltgpu_mod
    integer :: stepIT
""")
    s = edit(s, """    call update_particles()

    !********************************************************
    !*                 PRINT OUTPUT TO FILE                 *
""", """    !B200 offload: one internal step of every particle on the device (asynchronous);
    !  update_particles() below is no longer called
    call gpu_step()

    !********************************************************
    !*                 PRINT OUTPUT TO FILE                 *
""")
    # printOutput: fetch before writing, reset the device hit counters where the host's are reset
    s = edit(s, """    ! increment file number
    prcount = prcount + 1
""", """    ! increment file number
    prcount = prcount + 1

    !B200 offload: particle state of this print interval -> par, P_Salt, P_Temp, hit counters
    call gpu_fetch()
""")
    s = edit(s, """      !Reset Collision counters
      hitBottom = 0
      hitLand = 0
""", """      !Reset Collision counters
      hitBottom = 0
      hitLand = 0
      CALL ltgpu_check( ltgpu_reset_hits(gpu_ctx), 'ltgpu_reset_hits' )
""")
    s = edit(s, """    use param_mod,   only: numpar,SaltTempOn,TrackCollisions,WriteModelTiming
    use convert_mod, only: x2lon,y2lat
    integer :: n
""", """    use param_mod,   only: numpar,SaltTempOn,TrackCollisions,WriteModelTiming
    use convert_mod, only: x2lon,y2lat
    use ltgpu_mod
    integer :: n
""")
    s = edit(s, """      statuses(n) = getStatus(n)
""", """      statuses(n) = gpu_status(n)     !B200 offload: getStatus as fetched from the device
""")
    # fin_LTRANS: last fetch, statuses from the device, destroy
    s = edit(s, """    !Write final positions and status to endfile.csv
    IF(writeCSV)THEN
""", """    !B200 offload: final state, remaining ErrorLog lines
    call gpu_fetch()

    !Write final positions and status to endfile.csv
    IF(writeCSV)THEN
""")
    s = edit(s, """          par(n,pStatus) = getStatus(n)
          if(settlementon)then
""", """          par(n,pStatus) = gpu_status(n)
          if(settlementon)then
""")
    s = edit(s, """    !DEALLOCATE LOCAL VARIABLES
    DEALLOCATE(par)
    IF(ALLOCATED(hitBottom)) DEALLOCATE(hitBottom)
""", """    CALL gpu_fin()

    !DEALLOCATE LOCAL VARIABLES
    DEALLOCATE(par)
    IF(ALLOCATED(hitBottom)) DEALLOCATE(hitBottom)
""")
    # the four glue routines, as internal procedures of PROGRAM main (they see par, p, it, ix)
    s = edit(s, """end program""", """
  !*********************************************************
  !*              B200 offload (LTGPU_MOD)                 *
  !*********************************************************

  subroutine gpu_init()
    !Called from ini_LTRANS after initHydro: everything update_particles reads goes to the device
    use param_mod,      only: numpar,settlementon
    use hydro_mod,      only: gpuUploadGrid,gpuSetParticles,gpuPushHydro
    use boundary_mod,   only: gpuUploadBounds
    use settlement_mod, only: gpuUploadHabitat
    use ltgpu_mod
    type(ltgpu_params) :: prm
    integer(c_int64_t) :: counts(5),bad
    integer(c_int) :: ierr,nev
    type(ltgpu_event) :: ev(4096)

    call ltgpu_fill_params(prm)
    ierr = ltgpu_create(prm, 0_c_int, gpu_ctx)
    if(ierr == LTGPU_E_NODEVICE)then
      write(*,*) 'ERROR: no usable CUDA device; the particle loop has no CPU fallback'
      write(*,*) 'The Program Cannot Continue and Will Terminate'
      stop
    endif
    call ltgpu_check(ierr,'ltgpu_create')
    call gpuUploadGrid()
    call gpuUploadBounds()
    if(settlementon) call gpuUploadHabitat()
    call gpuSetParticles(par(:,pX),par(:,pY),par(:,pZ),par(:,pDOB))

    !the start-up screen (outside main bounds / inside an island / in no element) was done and
    !  logged by the serial code above; the device repeats it to set the same die / setOut flags
    ierr = ltgpu_screen_initial(gpu_ctx,counts,bad)
    if(ierr /= LTGPU_E_PARTICLE) call ltgpu_check(ierr,'ltgpu_screen_initial')
    do                              !its events duplicate lines already in ErrorLog.txt
      ierr = ltgpu_drain_events(gpu_ctx,ev,4096_c_int,nev)
      if(nev < 4096) exit
    enddo

    call gpuPushHydro(1)            !back, centre, forward records read by initHydro
    call gpuPushHydro(2)
    call gpuPushHydro(3)
    allocate(gpu_status(numpar))
    gpu_status = 0
  end subroutine gpu_init


  subroutine gpu_step()
    !Replaces update_particles(): internal step `it` of external step `p`.  With ErrorFlag
    !  outside 1..3 the reference STOPs at the first bad particle: checked at every step.
    use param_mod, only: ErrorFlag
    use ltgpu_mod
    integer(c_int) :: ierr,bad
    call ltgpu_check( ltgpu_step(gpu_ctx,int(p,c_int),int(it,c_int)), 'ltgpu_step' )
    if(ErrorFlag < 1 .OR. ErrorFlag > 3)then
      ierr = ltgpu_sync(gpu_ctx,bad)
      if(ierr == LTGPU_E_PARTICLE)then
        call ltgpu_write_errorlog()
        write(*,*) 'Particle ',bad,' met an error condition after ',int(ix(3)),' seconds'
        write(*,*) ' '
        write(*,*) 'The Program Cannot Continue and Will Terminate'
        stop
      endif
      call ltgpu_check(ierr,'ltgpu_sync')
    endif
  end subroutine gpu_step


  subroutine gpu_fetch()
    !Particle state -> the host arrays the writers use (printOutput, writeOutput, fin_LTRANS)
    use param_mod, only: SaltTempOn,TrackCollisions,settlementon
    use ltgpu_mod
    double precision, allocatable, target :: gx(:),gy(:),gz(:),ga(:),gl(:)
    type(c_ptr) :: ps,pt,phb,phl,pe
    integer :: np
    np = size(par,1)
    allocate(gx(np),gy(np),gz(np),ga(np),gl(np))
    ps = C_NULL_PTR; pt = C_NULL_PTR; phb = C_NULL_PTR; phl = C_NULL_PTR; pe = C_NULL_PTR
    if(SaltTempOn)then
      ps = C_LOC(P_Salt); pt = C_LOC(P_Temp)
    endif
    if(TrackCollisions)then
      phb = C_LOC(hitBottom); phl = C_LOC(hitLand)
    endif
    if(settlementon) pe = C_LOC(endpoly)
    call ltgpu_check( ltgpu_fetch(gpu_ctx, C_LOC(gx),C_LOC(gy),C_LOC(gz),C_LOC(ga),            &
         C_LOC(gpu_status), ps,pt, phb,phl, pe, C_LOC(gl),                                     &
         C_NULL_PTR,C_NULL_PTR,C_NULL_PTR), 'ltgpu_fetch' )
    par(:,pX) = gx
    par(:,pY) = gy
    par(:,pZ) = gz
    par(:,pAge) = ga
    par(:,pLifespan) = gl
    deallocate(gx,gy,gz,ga,gl)
    call ltgpu_write_errorlog()
  end subroutine gpu_fetch


  subroutine gpu_fin()
    use ltgpu_mod
    integer(c_int) :: ierr
    ierr = ltgpu_destroy(gpu_ctx)
    gpu_ctx = C_NULL_PTR
    if(allocated(gpu_status)) deallocate(gpu_status)
  end subroutine gpu_fin

end program""")
    # P_Salt, P_Temp, hit counters, endpoly are C_LOC targets
    s = edit(s, """  DOUBLE PRECISION, ALLOCATABLE, DIMENSION( : ) :: P_Salt,P_Temp
  INTEGER, ALLOCATABLE, DIMENSION(:) :: startpoly,endpoly,hitBottom,hitLand
""", """  DOUBLE PRECISION, ALLOCATABLE, TARGET, DIMENSION( : ) :: P_Salt,P_Temp
  INTEGER, ALLOCATABLE, TARGET, DIMENSION(:) :: startpoly,endpoly,hitBottom,hitLand
""")
    return s


# -------------------------------------------------------------------------------- makefile
def makefile(s):
    s = edit(s, """OBJS          = gridcell_module.o interpolation_module.o parameter_module.o \\
""", """# B200 offload: directory that holds libltrans_b200.so (ltransv.2b_b200/csrc of the library's tree)
LTGPU_LIBDIR  ?= .

OBJS          = gridcell_module.o interpolation_module.o parameter_module.o \\
""")
    s = edit(s, """%.o: %.f90
""", """# LTRANS with the particle loop on the GPU: same objects plus the ISO_C_BINDING interface module,
# linked against libltrans_b200.so (no CPU fallback: the executable needs a CUDA device).
# ltgpu_mod.o comes right after parameter_module.o (it USEs PARAM_MOD; HYDRO_MOD and the other
# patched modules USE LTGPU_MOD).
GPU_OBJS      = gridcell_module.o interpolation_module.o parameter_module.o ltgpu_mod.o \\
				point_in_polygon_module.o random_module.o tension_module.o  \\
				conversion_module.o hydrodynamic_module.o norm_module.o     \\
				boundary_module.o hor_turb_module.o settlement_module.o     \\
				ver_turb_module.o behavior_module.o

LTRANS_gpu : $(GPU_OBJS)
	@echo "  Compiling LTRANS.f90 (B200 offload)"
	@$(FC) $(FFLAGS) -o LTRANS_gpu.exe LTRANS.f90 $(GPU_OBJS) $(LIBS) \\
		-L$(LTGPU_LIBDIR) -lltrans_b200 -Wl,-rpath,$(LTGPU_LIBDIR)
	@\\rm *.o *.mod
	@echo "  "
	@echo "  Compilation Successfully Completed"
	@echo "  "

%.o: %.f90
""")
    return s


FILES = (("Model/hydrodynamic_module.f90", hydro), ("Model/boundary_module.f90", boundary),
         ("Model/settlement_module.f90", settlement), ("Model/LTRANS.f90", ltrans), ("Model/makefile", makefile))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(HERE, "patches")
    os.makedirs(out, exist_ok=True)
    for rel, fn in FILES:
        a = open(os.path.join(ref, rel), newline="", encoding="latin-1").read()
        crlf = "\r\n" in a
        b = fn(a.replace("\r\n", "\n"))
        if crlf:
            b = b.replace("\n", "\r\n")
        diff = difflib.unified_diff(a.splitlines(True), b.splitlines(True), "a/" + rel, "b/" + rel, n=3)
        name = os.path.join(out, os.path.basename(rel) + ".patch")
        with open(name, "w", newline="", encoding="latin-1") as f:
            f.writelines(diff)
        print("wrote", name)


if __name__ == "__main__":
    main()
