"""File adapters next to the hot path (sclmd/myio.py:80-135, 319-366).

The reference reads its electron-phonon coupling matrices (eta, xim, xip, zeta1, zeta2, ...) from NetCDF-4 files through
the `netCDF4` package.  NetCDF-4 is HDF5 underneath; the files the reference ships
(examples/current-induced/grapheneLambda-r-0.3-ver2.nc) hold plain contiguous, uncompressed float64 datasets, so a small
walker over the HDF5 structures is enough and no HDF5 library is needed:

  superblock (v0)  ->  root object header (v2, "OHDR")  ->  link-info message  ->  fractal heap direct block ("FHDB")
  ->  one link message per variable (name, object-header address)  ->  dataspace / datatype / layout messages.

Only what those files use is supported (version-2 object headers, links in one direct block or compact in the header,
contiguous layout, little-endian IEEE floats and integers); anything else raises `ValueError` naming the feature."""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


def _messages(buf, start, end, with_order):
    """(type, payload offset, size) of the header messages stored in buf[start:end]"""
    out = []
    pos = start
    while pos + 4 <= end:
        mtype = buf[pos]
        size = struct.unpack_from('<H', buf, pos + 1)[0]
        pos += 4 + (2 if with_order else 0)
        if pos + size > end:
            break
        out.append((mtype, pos, size))
        pos += size
    return out


def _object_header(buf, addr):
    """all messages of the version-2 object header at `addr`, continuation blocks followed"""
    if buf[addr:addr + 4] != b'OHDR' or buf[addr + 4] != 2:
        raise ValueError("HDF5: only version-2 object headers are supported (offset %d)" % addr)
    flags = buf[addr + 5]
    pos = addr + 6
    if flags & 0x20:
        pos += 16                      # access / modification / change / birth times
    if flags & 0x10:
        pos += 4                       # attribute storage phase-change values
    nsz = 1 << (flags & 3)
    chunk0 = int.from_bytes(buf[pos:pos + nsz], 'little')
    pos += nsz
    with_order = bool(flags & 0x04)
    msgs = _messages(buf, pos, pos + chunk0, with_order)
    todo = [m for m in msgs if m[0] == 0x10]
    while todo:
        _, p, _ = todo.pop()
        off, length = struct.unpack_from('<QQ', buf, p)
        if buf[off:off + 4] != b'OCHK':
            raise ValueError("HDF5: bad object-header continuation block at %d" % off)
        more = _messages(buf, off + 4, off + length - 4, with_order)
        msgs += more
        todo += [m for m in more if m[0] == 0x10]
    return msgs


def _link(buf, pos):
    """parse one link message at pos -> (name, address, next position) or None"""
    if buf[pos] != 1:
        return None
    flags = buf[pos + 1]
    p = pos + 2
    ltype = 0
    if flags & 0x08:
        ltype = buf[p]
        p += 1
    if flags & 0x04:
        p += 8                         # creation order
    if flags & 0x10:
        p += 1                         # character set
    nsz = 1 << (flags & 3)
    nlen = int.from_bytes(buf[p:p + nsz], 'little')
    p += nsz
    name = buf[p:p + nlen].decode('utf-8', 'replace')
    p += nlen
    if ltype != 0:
        raise ValueError("HDF5: soft / external links are not supported (%s)" % name)
    addr = struct.unpack_from('<Q', buf, p)[0]
    return name, addr, p + 8


def _links(buf, root):
    """name -> object header address of every link of the group whose object header is at `root`"""
    out = {}
    heap_addr = None
    for mtype, p, size in _object_header(buf, root):
        if mtype == 0x06:              # compact storage: the link message sits in the header itself
            lk = _link(buf, p)
            if lk:
                out[lk[0]] = lk[1]
        elif mtype == 0x02:            # link info -> fractal heap
            flags = buf[p + 1]
            q = p + 2 + (8 if flags & 1 else 0)
            heap_addr = struct.unpack_from('<Q', buf, q)[0]
    if heap_addr not in (None, _UNDEF):
        if buf[heap_addr:heap_addr + 4] != b'FRHP':
            raise ValueError("HDF5: fractal heap header expected at %d" % heap_addr)
        h = heap_addr
        hflags = buf[h + 9]
        # fractal heap header, version 0 (HDF5 file format specification III.G): signature 4, version 1, heap-id length 2,
        # filter length 2, flags 1, max managed object size 4, twelve 8-byte counters/addresses, table width 2,
        # starting block size 8, max direct block size 8, then the fields used here
        max_heap_bits = struct.unpack_from('<H', buf, h + 128)[0]
        root_block = struct.unpack_from('<Q', buf, h + 132)[0]
        cur_rows = struct.unpack_from('<H', buf, h + 140)[0]
        if cur_rows != 0:
            raise ValueError("HDF5: fractal heaps with indirect blocks are not supported")
        if buf[root_block:root_block + 4] != b'FHDB':
            raise ValueError("HDF5: fractal heap direct block expected at %d" % root_block)
        pos = root_block + 4 + 1 + 8 + (max_heap_bits + 7) // 8 + (4 if hflags & 2 else 0)
        while pos < len(buf):
            lk = _link(buf, pos)
            if lk is None:
                break
            out[lk[0]] = lk[1]
            pos = lk[2]
    return out


def _dataset(buf, addr):
    """ndarray of the dataset whose object header is at addr, or None if it is not a plain contiguous numeric dataset"""
    dims = dtype = data = None
    for mtype, p, size in _object_header(buf, addr):
        if mtype == 0x01:
            ver, rank, fl = buf[p], buf[p + 1], buf[p + 2]
            q = p + (8 if ver == 1 else 4)
            dims = struct.unpack_from('<%dQ' % rank, buf, q) if rank else ()
        elif mtype == 0x03:
            cls = buf[p] & 0x0F
            bits0 = buf[p + 1]
            nbytes = struct.unpack_from('<I', buf, p + 4)[0]
            order = '>' if bits0 & 1 else '<'
            if cls == 1:
                dtype = np.dtype(order + 'f%d' % nbytes)
            elif cls == 0:
                dtype = np.dtype(order + ('i' if bits0 & 0x08 else 'u') + '%d' % nbytes)
        elif mtype == 0x08:
            if buf[p] != 3:
                raise ValueError("HDF5: data-layout message version %d is not supported" % buf[p])
            if buf[p + 1] == 1:        # contiguous
                data = struct.unpack_from('<QQ', buf, p + 2)
    if dims is None or dtype is None or data is None or data[0] == _UNDEF:
        return None
    n = int(np.prod(dims)) if dims else 1
    if n * dtype.itemsize > data[1]:
        return None
    return np.frombuffer(buf, dtype=dtype, count=n, offset=data[0]).reshape(dims).astype(dtype.newbyteorder('='))


def read_nc_variables(filename):
    """dict name -> ndarray of every numeric variable of a NetCDF file: NetCDF-4 / HDF5 (contiguous datasets of the root group,
    the walker above) or NetCDF classic (scipy.io.netcdf_file)"""
    with open(filename, 'rb') as f:
        buf = f.read()
    if buf[:3] == b'CDF':
        from scipy.io import netcdf_file
        with netcdf_file(filename, 'r', mmap=False) as nf:
            return {k: np.array(v[:], dtype=float if v.typecode() in 'fd' else None) for k, v in nf.variables.items()}
    if buf[:8] != b'\x89HDF\r\n\x1a\n':
        raise ValueError("%s is not an HDF5 / NetCDF-4 file" % filename)
    ver = buf[8]
    if ver in (0, 1):
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("HDF5: only 8-byte offsets and lengths are supported")
        root = struct.unpack_from('<Q', buf, 24 + 32 + (4 if ver == 1 else 0) + 8)[0]
    elif ver in (2, 3):
        root = struct.unpack_from('<Q', buf, 12 + 8 + 8 + 8)[0]
    else:
        raise ValueError("HDF5: superblock version %d is not supported" % ver)
    out = {}
    for name, addr in _links(buf, root).items():
        try:
            arr = _dataset(buf, addr)
        except ValueError:
            arr = None
        if arr is not None:
            out[name] = arr
    return out


class _Dataset:
    """the part of netCDF4.Dataset the reference's scripts use for reading: ds['name'][:] and ds.variables['name']"""

    def __init__(self, filename, mode='r'):
        if mode != 'r':
            raise ValueError("sclmd_b200.myio.Dataset is read-only (checkpoints are written by md.dump as NetCDF classic files with scipy.io.netcdf_file)")
        self.variables = read_nc_variables(filename)

    def __getitem__(self, name):
        return self.variables[name]

    def close(self):
        pass


Dataset = _Dataset


def ReadNetCDFVar(file, var):
    """sclmd/myio.py:184-189, md.py ReadNetCDFVar"""
    return read_nc_variables(file)[var]


def Write2NetCDFFile(file, var, varLabel, dimensions, units=None, description=None):
    """sclmd/myio.py:174-181 on an open scipy.io.netcdf_file (or netCDF4.Dataset): one float64 variable"""
    tmp = file.createVariable(varLabel, 'd', dimensions)
    tmp[:] = np.asarray(var, dtype=float)
    if units:
        tmp.units = units
    if description:
        tmp.description = description


def ReadMDNCFile(filename):
    """sclmd/myio.py:192-211: cell, coordinates and atom lists of an MD set-up file"""
    v = read_nc_variables(filename)

    class mdmath:
        pass

    mdmath.filename = filename
    mdmath.cell, mdmath.xyz = v['UnitCell'], v['XYZ']
    mdmath.dynatom, mdmath.atomlist = v['DynamicAtoms'], v['AtomList']
    return mdmath


def ord2idx(order):
    """sclmd/myio.py:291-297: dof permutation of a 1-based atom order"""
    order = np.asarray(order, dtype=int)
    return (3 * (order[:, None] - 1) + np.arange(3)[None, :]).reshape(-1)


def ReadDynmat(filename, order=None):
    """sclmd/myio.py:214-250: dynamical matrix in real space from the phonon file of Inelastica's PHrun (hw, U, DynamicAtoms):
    dyn = U^T diag(hw^2) U restricted to the dynamic atoms, symmetrised; `order` = new 1-based atom order"""
    v = read_nc_variables(filename)
    hw, fullU, dynatoms = v['hw'], v['U'], v['DynamicAtoms']
    first, last = int(dynatoms[0]) - 1, int(dynatoms[-1])
    U = np.array([np.asarray(row)[first:last].reshape(-1) for row in fullU], dtype=float)
    if order is not None:
        if 3 * len(order) != len(hw):
            raise ValueError("ReadDynmat: length of order error!")
        U = U[:, ord2idx(order)]
    dyn = (U.T * hw ** 2) @ U
    return 0.5 * (dyn + dyn.T), U, hw


def ReadSig(filename):
    """sclmd/myio.py:300-316: frequency grid and the two lead self-energies"""
    v = read_nc_variables(filename)

    class eph:
        pass

    eph.wl = v['Wlist']
    eph.SigL = v['ReSigL'] + 1j * v['ImSigL']
    eph.SigR = v['ReSigR'] + 1j * v['ImSigR']
    return eph


def ReadwbLambda(filename, order=None):
    """sclmd/myio.py:319-336: wide-band electron-phonon matrices (bias is zero by construction)"""
    v = read_nc_variables(filename)
    return 0.0, v['eta'], v['xim'], v['xip'], v['zeta1'], v['zeta2']


def ReadLambda(filename, w0, order=None):
    """sclmd/myio.py:339-366: friction / non-conservative / renormalisation / Berry matrices at the grid energy nearest to w0,
    from the energy-resolved Pi and Lambda of a biased junction"""
    from .functions import nearest
    v = read_nc_variables(filename)
    wl, mus = v['wl'], v['muLR']
    bias = mus[0] - mus[1]
    k = nearest(w0, wl)
    w00 = wl[k]
    im, re, lam = v['ImPir2'][k], v['RePir2'][k], v['ReLamLR'][k]
    eta = -(im + im.T) / 2 / w00
    zeta2 = -(im - im.T) / 2 / w00 / bias
    xim = -(re - re.T) / 2 / bias
    zeta1 = (re + re.T) / 2 / bias
    xip = -np.pi * (lam + lam.T) / 2 / w00
    return bias, eta, xim, xip, zeta1, zeta2


def WriteEPHNCfile(filename, wl, hw, U, DynMat, SigL, SigR, Friction, NC, NCP, zeta1, zeta2):
    """sclmd/myio.py:138-171 in NetCDF classic format (same dimension and variable names; netCDF4 reads it)"""
    from scipy.io import netcdf_file
    SigL, SigR = np.asarray(SigL), np.asarray(SigR)
    f = netcdf_file(filename, 'w')
    f.createDimension('NPh', len(hw))
    f.createDimension('NWl', len(wl))
    f.createDimension('Nsl', len(SigL[0]))
    f.createDimension('Nsr', len(SigR[0]))

    def put(name, arr, dims, units):
        v = f.createVariable(name, 'd', dims)
        v[:] = np.asarray(arr, dtype=float)
        v.units = units
    put('Wlist', wl, ('NWl',), 'eV')
    put('hw', hw, ('NPh',), 'eV')
    put('U', U, ('NPh', 'NPh'), 'None')
    put('DynMat', DynMat, ('NPh', 'NPh'), 'eV**2')
    put('ReSigL', SigL.real, ('NWl', 'Nsl', 'Nsl'), 'eV**2')
    put('ImSigL', SigL.imag, ('NWl', 'Nsl', 'Nsl'), 'eV**2')
    put('ReSigR', SigR.real, ('NWl', 'Nsr', 'Nsr'), 'eV**2')
    put('ImSigR', SigR.imag, ('NWl', 'Nsr', 'Nsr'), 'eV**2')
    for name, arr in (('Friction', Friction), ('NC', NC), ('NCP', NCP), ('zeta1', zeta1), ('zeta2', zeta2)):
        put(name, arr, ('NPh', 'NPh'), 'eV**2')
    f.close()


def ReadEPHNCFile(filename):
    """sclmd/myio.py:80-106"""
    e = ReadNewEPHNCFile(filename)
    return e


def ReadNewEPHNCFile(filename):
    """sclmd/myio.py:109-135: dynamical matrix, lead self-energies and the electron friction / non-conservative / Berry /
    renormalisation matrices of a junction"""
    v = read_nc_variables(filename)

    class eph:
        pass

    eph.filename = filename
    eph.wl, eph.hw, eph.U, eph.DynMat = v['Wlist'], v['hw'], v['U'], v['DynMat']
    eph.SigL = v['ReSigL'] + 1j * v['ImSigL']
    eph.SigR = v['ReSigR'] + 1j * v['ImSigR']
    eph.efric, eph.xim, eph.xip = v['Friction'], v['NC'], v['NCP']
    eph.zeta1, eph.zeta2 = v.get('zeta1'), v.get('zeta2')
    return eph


def reordxyz(anr, xyz, ord):
    """sclmd/myio.py:64-77: move the atoms listed in `ord` (1-based, a contiguous index range in any order) into that order;
    atoms before and after the range keep their places.  Returns the permuted atom numbers and coordinates."""
    lo, hi = min(ord), max(ord)
    perm = list(range(lo - 1)) + [i - 1 for i in ord] + list(range(hi, len(xyz)))
    if len(perm) != len(anr):
        raise ValueError("reordxyz:length error")
    return [anr[i] for i in perm], [xyz[i] for i in perm]
