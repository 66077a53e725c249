"""SHDOM grid data structure (host side, numpy): base grid and cell splitting.

Integer arrays are bit-exact restatements of the reference formulas (paths relative to the AT3D
checkout):

* ``new_grids``            -- NEW_GRIDS, src/polarized/shdomsub2.f:16-81
* ``init_cell_structure``  -- INIT_CELL_STRUCTURE, src/polarized/shdomsub2.f:84-309
* ``boundary_pnts``        -- BOUNDARY_PNTS, src/polarized/shdomsub1.f:2173-2215
* ``divide_cell``          -- DIVIDE_CELL / MATCH_NEIGHBOR_FACE / INHERIT_NEIGHBOR /
                              NEW_GRID_POINTS / MATCH_GRID_POINT, shdomsub1.f:5286-5696
* ``sh_sizes``, ``lofj``   -- ML/MM/NLM bookkeeping, at3d/solver.py:2188-2191,
                              shdomsub1.f:1057-1065

* ``make_grid``            -- at3d/grid.py:36-97, the property-grid dataset (``x, y, z, delx, dely``) every
                              scatterer and sensor bounding box is defined on

All arrays use the reference layout: Fortran order, 1-based index contents.
"""
import numpy as np
from ._dataset import Dataset


def make_grid(delx, npx, dely, npy, z, nx=None, ny=None, nz=None):
    """Regular horizontal grid starting at 0 with spacings `delx`, `dely` and the (irregular) vertical levels `z`; `nx`,
    `ny`, `nz` optionally ask for a different base-grid resolution of the solver (at3d/grid.py:36-97)."""
    z = np.asarray(z)
    if (z.ndim != 1) or (z.size < 2) or (not np.all(np.sort(z) == z)) or (not np.all(z >= 0.0)) or \
            (np.unique(z).size != z.size):
        raise ValueError('z must be >= 0, strictly increasing, 1-D and contain at least 2 points.')
    grid = Dataset(x=np.linspace(0.0, delx * (npx - 1), npx), y=np.linspace(0.0, dely * (npy - 1), npy), z=z,
                   delx=delx, dely=dely)
    for name, value in (('nx', nx), ('ny', ny), ('nz', nz)):
        if value is not None:
            grid[name] = value
    return grid

OPPFACE = (2, 1, 4, 3, 6, 5)
GRIDFACE = ((1, 3, 5, 7), (2, 4, 6, 8), (1, 2, 5, 6), (3, 4, 7, 8), (1, 2, 3, 4), (5, 6, 7, 8))
# GRIDCORNER(2,4,3), FACEGRID(2,4,3) of NEW_GRID_POINTS (shdomsub1.f:5528-5531), [idir][i] -> pair
GRIDCORNER = (((1, 2), (3, 4), (5, 6), (7, 8)),
              ((1, 3), (2, 4), (5, 7), (6, 8)),
              ((1, 5), (2, 6), (3, 7), (4, 8)))
FACEGRID = (((3, 5), (4, 5), (3, 6), (4, 6)),
            ((1, 5), (2, 5), (1, 6), (2, 6)),
            ((1, 3), (2, 3), (1, 4), (2, 4)))


def btest(val, bit):
    return bool((int(val) >> bit) & 1)


def sh_sizes(nmu, nphi):
    """ML, MM, NLM, NLEG for an (NMU, NPHI) angle set (solver.py:2188-2191)."""
    ml = nmu - 1
    mm = max(0, nphi // 2 - 1)
    nlm = (2 * mm + 1) * (ml + 1) - mm * (mm + 1)
    return ml, mm, nlm


def lofj(ml, mm):
    """l index of each SH term j (shdomsub1.f:1057-1065)."""
    out = []
    for l in range(ml + 1):
        me = min(l, mm)
        out.extend([l] * (2 * me + 1))
    return np.asarray(out, dtype=np.int32)


def grid_sizes(nx, ny, nz, bcflag, ipflag):
    """NX1, NY1, number of base points and cells (solver.py:2121-2132)."""
    nx1, ny1 = nx + 1, ny + 1
    if (bcflag & 5) or btest(ipflag, 0):
        nx1 -= 1
    if (bcflag & 10) or btest(ipflag, 1):
        ny1 -= 1
    nxc = nx + (1 if btest(bcflag, 0) else 0) - (1 if btest(bcflag, 2) else 0)
    nyc = ny + (1 if btest(bcflag, 1) else 0) - (1 if btest(bcflag, 3) else 0)
    return nx1, ny1, nx1 * ny1 * nz, (nz - 1) * nxc * nyc


def new_grids(bcflag, gridtype, npx, npy, npz, nx, ny, nz, xstart, ystart, delxp, delyp, zlevels):
    """NEW_GRIDS: returns XGRID(NX+1), YGRID(NY+1), ZGRID(NZ) in float32 arithmetic."""
    f = np.float32
    delxp = f(delxp) if delxp > 0 else f(1.0)
    delyp = f(delyp) if delyp > 0 else f(1.0)
    xgrid = np.zeros(nx + 1, dtype=f)
    ygrid = np.zeros(ny + 1, dtype=f)
    zgrid = np.zeros(nz, dtype=f)
    zlevels = np.asarray(zlevels, dtype=f)
    if btest(bcflag, 2):
        for ix in range(1, nx + 1):
            xgrid[ix - 1] = f(xstart) + f(f(f(ix - 1) * f(delxp * f(npx - 1))) / f(nx - 1))
    else:
        for ix in range(1, nx + 2):
            xgrid[ix - 1] = f(xstart) + f(f(f(ix - 1) * f(delxp * f(npx))) / f(nx))
    if btest(bcflag, 3):
        for iy in range(1, ny + 1):
            ygrid[iy - 1] = f(ystart) + f(f(f(iy - 1) * f(delyp * f(npy - 1))) / f(ny - 1))
    else:
        for iy in range(1, ny + 2):
            ygrid[iy - 1] = f(ystart) + f(f(f(iy - 1) * f(delyp * f(npy))) / f(ny))
    if gridtype == 'P':
        if nz != npz:
            raise ValueError('NEW_GRIDS: must have NZ=NPZ for gridtype P')
        zgrid[:] = zlevels[:nz]
    elif gridtype == 'E':
        for iz in range(1, nz + 1):
            zgrid[iz - 1] = zlevels[0] + f(f(f(iz - 1) * f(zlevels[npz - 1] - zlevels[0])) / f(nz - 1))
    else:
        raise ValueError('NEW_GRIDS: Illegal grid type in Z')
    return xgrid, ygrid, zgrid


def init_cell_structure(bcflag, ipflag, nx, ny, nz, nx1, ny1, xgrid, ygrid, zgrid, maxic=None, maxig=None):
    """INIT_CELL_STRUCTURE.  Returns (npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags)."""
    npts = nx1 * ny1 * nz
    nxc = nx + (1 if btest(bcflag, 0) else 0)
    if btest(bcflag, 2) and not btest(ipflag, 0):
        nxc = nx - 1
    nyc = ny + (1 if btest(bcflag, 1) else 0)
    if btest(bcflag, 3) and not btest(ipflag, 1):
        nyc = ny - 1
    ncells = nxc * nyc * (nz - 1)
    maxig = max(maxig or 0, npts)
    maxic = max(maxic or 0, ncells)
    gridpos = np.zeros((3, maxig), dtype=np.float32, order='F')
    gx, gy, gz = np.meshgrid(np.asarray(xgrid[:nx1], np.float32), np.asarray(ygrid[:ny1], np.float32),
                             np.asarray(zgrid[:nz], np.float32), indexing='ij')
    gridpos[0, :npts] = gx.ravel()
    gridpos[1, :npts] = gy.ravel()
    gridpos[2, :npts] = gz.ravel()

    gridptr = np.zeros((8, maxic), dtype=np.int32, order='F')
    neighptr = np.zeros((6, maxic), dtype=np.int32, order='F')
    treeptr = np.zeros((2, maxic), dtype=np.int32, order='F')
    cellflags = np.zeros(maxic, dtype=np.int16)

    ix = np.arange(1, nxc + 1)
    iy = np.arange(1, nyc + 1)
    iz = np.arange(1, nz)
    IX, IY, IZ = np.meshgrid(ix, iy, iz, indexing='ij')
    IX = IX.ravel(); IY = IY.ravel(); IZ = IZ.ravel()
    I = np.arange(1, ncells + 1)
    if btest(ipflag, 0):
        ix0, ix1 = IX, IX
    elif btest(bcflag, 0):
        ix0, ix1 = np.maximum(1, IX - 1), np.minimum(nx, IX)
    else:
        ix0, ix1 = IX, IX + 1
    if btest(ipflag, 1):
        iy0, iy1 = IY, IY
    elif btest(bcflag, 1):
        iy0, iy1 = np.maximum(1, IY - 1), np.minimum(ny, IY)
    else:
        iy0, iy1 = IY, IY + 1
    for k, (zz, yy, xx) in enumerate([(IZ, iy0, ix0), (IZ, iy0, ix1), (IZ, iy1, ix0), (IZ, iy1, ix1),
                                      (IZ + 1, iy0, ix0), (IZ + 1, iy0, ix1), (IZ + 1, iy1, ix0),
                                      (IZ + 1, iy1, ix1)]):
        gridptr[k, :ncells] = zz + nz * (yy - 1) + nz * ny1 * (xx - 1)
    flags = np.zeros(ncells, dtype=np.int64)
    sx = (nz - 1) * nyc
    n1 = np.zeros(ncells, dtype=np.int64); n2 = np.zeros(ncells, dtype=np.int64)
    if btest(ipflag, 0):
        n1[:] = I; n2[:] = I; flags |= 1
    elif btest(bcflag, 0):
        n1[:] = I - sx; n2[:] = I + sx
        lo = IX == 1; hi = IX == nxc
        n1[lo] = I[lo]; flags[lo] |= 1
        n2[hi] = I[hi]; flags[hi] |= 1
    elif btest(bcflag, 2):
        n1[:] = I - sx; n2[:] = I + sx
        n1[IX == 1] = 0; n2[IX == nxc] = 0
    else:
        n1[:] = np.where(IX == 1, I + (nxc - 1) * sx, I - sx)
        n2[:] = np.where(IX == nx, I - (nxc - 1) * sx, I + sx)
    sy = nz - 1
    n3 = np.zeros(ncells, dtype=np.int64); n4 = np.zeros(ncells, dtype=np.int64)
    if btest(ipflag, 1):
        n3[:] = I; n4[:] = I; flags |= 2
    elif btest(bcflag, 1):
        n3[:] = I - sy; n4[:] = I + sy
        lo = IY == 1; hi = IY == nyc
        n3[lo] = I[lo]; flags[lo] |= 2
        n4[hi] = I[hi]; flags[hi] |= 2
    elif btest(bcflag, 3):
        n3[:] = I - sy; n4[:] = I + sy
        n3[IY == 1] = 0; n4[IY == nyc] = 0
    else:
        n3[:] = np.where(IY == 1, I + (nyc - 1) * sy, I - sy)
        n4[:] = np.where(IY == ny, I - (nyc - 1) * sy, I + sy)
    n5 = np.where(IZ == 1, 0, I - 1)
    n6 = np.where(IZ == nz - 1, 0, I + 1)
    for k, n in enumerate((n1, n2, n3, n4, n5, n6)):
        neighptr[k, :ncells] = n
    cellflags[:ncells] = flags
    return npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags


def boundary_pnts(npts, gridpos, zbot, ztop):
    """BOUNDARY_PNTS: sorted (1-based) lists of top and bottom boundary points -> bcptr[maxnbc,2]."""
    z = gridpos[2, :npts]
    top = np.nonzero(z >= np.float32(ztop))[0] + 1
    bot = np.nonzero(z <= np.float32(zbot))[0] + 1
    maxnbc = max(len(top), len(bot), 1)
    bcptr = np.zeros((maxnbc, 2), dtype=np.int32, order='F')
    bcptr[:len(top), 0] = top
    bcptr[:len(bot), 1] = bot
    return len(top), len(bot), bcptr


class CellTree:
    """Mutable view of the grid arrays for cell splitting (DIVIDE_CELL and helpers)."""

    def __init__(self, npts, ncells, gridpos, gridptr, neighptr, treeptr, cellflags):
        self.npts, self.ncells = npts, ncells
        self.gridpos, self.gridptr, self.neighptr = gridpos, gridptr, neighptr
        self.treeptr, self.cellflags = treeptr, cellflags

    def _grow(self, ncell_need, npt_need):
        if ncell_need > self.gridptr.shape[1]:
            n = max(ncell_need, 2 * self.gridptr.shape[1])
            for name in ('gridptr', 'neighptr', 'treeptr'):
                old = getattr(self, name)
                new = np.zeros((old.shape[0], n), dtype=old.dtype, order='F')
                new[:, :old.shape[1]] = old
                setattr(self, name, new)
            cf = np.zeros(n, dtype=np.int16)
            cf[:self.cellflags.shape[0]] = self.cellflags
            self.cellflags = cf
        if npt_need > self.gridpos.shape[1]:
            n = max(npt_need, 2 * self.gridpos.shape[1])
            new = np.zeros((3, n), dtype=np.float32, order='F')
            new[:, :self.gridpos.shape[1]] = self.gridpos
            self.gridpos = new

    # -- MATCH_GRID_POINT (shdomsub1.f:5599-5696)
    def _match_grid_point(self, xp, yp, zp, icell, iface):
        gp, gpos, tp, cf = self.gridptr, self.gridpos, self.treeptr, self.cellflags
        idir = (iface + 1) // 2
        kface = OPPFACE[iface - 1]
        ic = abs(int(self.neighptr[iface - 1, icell - 1]))
        if ic == 0:
            return 0
        stack = []
        while True:
            while tp[1, ic - 1] == 0:
                for i in range(4):
                    ipt = int(gp[GRIDFACE[kface - 1][i] - 1, ic - 1])
                    if xp == gpos[0, ipt - 1] and yp == gpos[1, ipt - 1] and zp == gpos[2, ipt - 1]:
                        return ipt
                if not stack:
                    return 0
                ic = stack.pop()
            d = (int(cf[ic - 1]) >> 2) & 3
            if d == 0:
                raise RuntimeError('MATCH_GRID_POINT: No split direction')
            ic1 = int(tp[1, ic - 1])
            if d == idir:
                ic = ic1 + 1 - ((iface - 1) % 2)
            else:
                ic = ic1
                p = (xp, yp, zp)[d - 1]
                split = gpos[d - 1, gp[7, ic1 - 1] - 1]
                if p == split:
                    stack.append(ic + 1)
                elif p > split:
                    ic = ic + 1

    # -- NEW_GRID_POINTS (shdomsub1.f:5509-5595)
    def _new_grid_points(self, idir, icell, newcell):
        gp, gpos = self.gridptr, self.gridpos
        gp[:, newcell - 1] = gp[:, icell - 1]
        gp[:, newcell] = gp[:, icell - 1]
        newpoints = np.zeros((3, 4), dtype=np.int32, order='F')
        f = np.float32
        for i in range(4):
            i1, i2 = GRIDCORNER[idir - 1][i]
            ip1, ip2 = int(gp[i1 - 1, icell - 1]), int(gp[i2 - 1, icell - 1])
            xp = f(f(gpos[0, ip1 - 1] + gpos[0, ip2 - 1]) / f(2))
            yp = f(f(gpos[1, ip1 - 1] + gpos[1, ip2 - 1]) / f(2))
            zp = f(f(gpos[2, ip1 - 1] + gpos[2, ip2 - 1]) / f(2))
            iface1, iface2 = FACEGRID[idir - 1][i]
            ipmatch = self._match_grid_point(xp, yp, zp, icell, iface1)
            if ipmatch == 0:
                ipmatch = self._match_grid_point(xp, yp, zp, icell, iface2)
            if ipmatch == 0:
                icell2 = abs(int(self.neighptr[iface1 - 1, icell - 1]))
                if icell2 > 0:
                    ipmatch = self._match_grid_point(xp, yp, zp, icell2, iface2)
            if ipmatch == 0:
                self.npts += 1
                self._grow(0, self.npts)
                gp, gpos = self.gridptr, self.gridpos
                gp[i2 - 1, newcell - 1] = self.npts
                gp[i1 - 1, newcell] = self.npts
                gpos[:, self.npts - 1] = (xp, yp, zp)
                newpoints[:, i] = (ip1, ip2, self.npts)
            else:
                gp[i2 - 1, newcell - 1] = ipmatch
                gp[i1 - 1, newcell] = ipmatch
                newpoints[2, i] = 0
        return newpoints

    # -- INHERIT_NEIGHBOR (shdomsub1.f:5454-5501)
    def _inherit_neighbor(self, icell, iface, inn):
        nb, tp, cf = self.neighptr, self.treeptr, self.cellflags
        jface = (iface + 1) // 2
        ic = icell
        stack = []
        while True:
            nb[iface - 1, ic - 1] = inn
            if tp[1, ic - 1] == 0:
                if not stack:
                    return
                ic = stack.pop()
            else:
                d = (int(cf[ic - 1]) >> 2) & 3
                if d == jface:
                    ic = int(tp[1, ic - 1]) + ((iface - 1) % 2)
                else:
                    stack.append(int(tp[1, ic - 1]) + 1)
                    ic = int(tp[1, ic - 1])

    # -- MATCH_NEIGHBOR_FACE (shdomsub1.f:5371-5450)
    def _match_neighbor_face(self, iface, ic):
        nb, tp, cf, gp, gpos = self.neighptr, self.treeptr, self.cellflags, self.gridptr, self.gridpos
        inn = abs(int(nb[iface - 1, ic - 1]))
        if inn == 0:
            return
        jface = (iface + 1) // 2
        ic1, ic8 = int(gp[0, ic - 1]), int(gp[7, ic - 1])
        dir1 = (jface - 1 + 1) % 3 + 1
        dir2 = (jface - 1 + 2) % 3 + 1
        f = np.float32
        pos = [f(0), f(0), f(0)]
        pos[dir1 - 1] = f(f(gpos[dir1 - 1, ic1 - 1] + gpos[dir1 - 1, ic8 - 1]) / f(2))
        pos[dir2 - 1] = f(f(gpos[dir2 - 1, ic1 - 1] + gpos[dir2 - 1, ic8 - 1]) / f(2))
        done = False
        while not done and ic != inn:
            in1, in8 = int(gp[0, inn - 1]), int(gp[7, inn - 1])
            if (gpos[dir1 - 1, ic1 - 1] >= gpos[dir1 - 1, in1 - 1] and
                    gpos[dir1 - 1, ic8 - 1] <= gpos[dir1 - 1, in8 - 1] and
                    gpos[dir2 - 1, ic1 - 1] >= gpos[dir2 - 1, in1 - 1] and
                    gpos[dir2 - 1, ic8 - 1] <= gpos[dir2 - 1, in8 - 1]):
                nb[iface - 1, ic - 1] = inn if tp[1, inn - 1] == 0 else -inn
            of = OPPFACE[iface - 1]
            if (gpos[dir1 - 1, in1 - 1] >= gpos[dir1 - 1, ic1 - 1] and
                    gpos[dir1 - 1, in8 - 1] <= gpos[dir1 - 1, ic8 - 1] and
                    gpos[dir2 - 1, in1 - 1] >= gpos[dir2 - 1, ic1 - 1] and
                    gpos[dir2 - 1, in8 - 1] <= gpos[dir2 - 1, ic8 - 1]):
                self._inherit_neighbor(inn, of, ic)
            else:
                nb[of - 1, inn - 1] = -abs(int(nb[of - 1, inn - 1]))
            if tp[1, inn - 1] == 0:
                done = True
            else:
                d = (int(cf[inn - 1]) >> 2) & 3
                inn2 = int(tp[1, inn - 1])
                if d == jface:
                    inn = inn2 + 1 - ((iface - 1) % 2)
                else:
                    inn = inn2 + 1 if pos[d - 1] > gpos[d - 1, gp[7, inn2 - 1] - 1] else inn2

    # -- DIVIDE_CELL (shdomsub1.f:5286-5365); idir 1=X, 2=Y, 3=Z
    def divide_cell(self, icell, idir):
        if self.treeptr[1, icell - 1] != 0:
            raise RuntimeError('DIVIDE_CELL: Cannot divide already split cell.')
        newcell = self.ncells + 1
        self.ncells += 2
        self._grow(self.ncells, self.npts + 4)
        tp, cf, nb = self.treeptr, self.cellflags, self.neighptr
        tp[1, icell - 1] = newcell
        tp[:, newcell - 1] = (icell, 0)
        tp[:, newcell] = (icell, 0)
        cf[icell - 1] = int(cf[icell - 1]) | (idir << 2)
        cf[newcell - 1] = int(cf[icell - 1]) & 3
        cf[newcell] = int(cf[icell - 1]) & 3
        newpoints = self._new_grid_points(idir, icell, newcell)
        nb = self.neighptr
        for iface in range(1, 7):
            if nb[iface - 1, icell - 1] == icell:
                nb[iface - 1, newcell - 1] = newcell
                nb[iface - 1, newcell] = newcell + 1
            elif iface == 2 * idir:
                nb[iface - 1, newcell - 1] = newcell + 1
                nb[iface - 1, newcell] = nb[iface - 1, icell - 1]
            elif iface == 2 * idir - 1:
                nb[iface - 1, newcell] = newcell
                nb[iface - 1, newcell - 1] = nb[iface - 1, icell - 1]
            else:
                nb[iface - 1, newcell - 1] = nb[iface - 1, icell - 1]
                nb[iface - 1, newcell] = nb[iface - 1, icell - 1]
            self._match_neighbor_face(iface, newcell)
            self._match_neighbor_face(iface, newcell + 1)
        return newpoints
