"""VTU output and restart for the device-resident solution (the data format either side of the Newton path).

Mirrors what DuMux's `VtkOutputModule` writes for cell-centred models (dumux/io/vtkoutputmodule.hh:  one <Piece> of an
UnstructuredGrid, cell data as Float32 ASCII arrays in element order, x fastest, plus the `process rank` field) and what
`loadSolution` reads back for a restart (dumux/io/loadsolution.hh:332: the primary variables are looked up among the cell
data by `IOFields::primaryVariableName`, e.g. p_aq and S_napl for the p0-s1 2p model, porousmediumflow/2p/iofields.hh:52-61).
The grid is the structured box of `dmx_grid_structured` / `dmx_grid_tensor`: quadrilaterals / hexahedra (VTK types 9 / 12),
lines in 1-D (type 3), vertices in lexicographic order as YaspGrid numbers them.
Host-side control code: the fields come from `Engine.output_fields()` (device kernel) or, in tests, from the oracle.
"""
import xml.etree.ElementTree as ET

import numpy as np

_VTK_TYPE = {1: 3, 2: 9, 3: 12}
# VTK corner order of a quad / hexahedron in terms of the lexicographic (YaspGrid) corner index
_CORNERS = {1: [0, 1], 2: [0, 1, 3, 2], 3: [0, 1, 3, 2, 4, 5, 7, 6]}


def _fmt(a):
    """Float32 ASCII, 12 values per line like Dune's VTKWriter (ascii output type)."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    lines = [" ".join(f"{v:.6g}" for v in a[i:i + 12]) for i in range(0, a.size, 12)]
    return "\n          " + "\n          ".join(lines) + "\n        "


def write_vtu(path, node_coords, fields, rank=0):
    """node_coords: per-axis node coordinate arrays (len cells+1); fields: ordered mapping name -> array[n] or array[n, ncomp]
    (cell data, x fastest).  Adds the `process rank` field the reference writes."""
    dim = len(node_coords)
    nn = [len(c) for c in node_coords]
    nc = [k - 1 for k in nn]
    n = int(np.prod(nc))
    grids = np.meshgrid(*[np.asarray(c, dtype=np.float64) for c in node_coords], indexing="ij")
    pts = np.zeros((int(np.prod(nn)), 3))
    for a in range(dim):
        pts[:, a] = np.transpose(grids[a], axes=range(dim - 1, -1, -1)).reshape(-1)      # x fastest
    # cell -> node connectivity
    idx = np.indices(nc[::-1]).reshape(dim, -1)[::-1]                                   # idx[a] = cell coordinate along axis a, x fastest
    stride = [int(np.prod(nn[:a])) for a in range(dim)]
    base = sum(idx[a] * stride[a] for a in range(dim))
    lex = []
    for corner in range(2 ** dim):
        lex.append(base + sum(((corner >> a) & 1) * stride[a] for a in range(dim)))
    conn = np.stack([lex[c] for c in _CORNERS[dim]], axis=1)
    root = ET.Element("VTKFile", type="UnstructuredGrid", version="0.1", byte_order="LittleEndian")
    piece = ET.SubElement(ET.SubElement(root, "UnstructuredGrid"), "Piece", NumberOfCells=str(n), NumberOfPoints=str(pts.shape[0]))
    names = list(fields.keys())
    cd = ET.SubElement(piece, "CellData", Scalars=names[0] if names else "process rank")
    for name in names:
        v = np.asarray(fields[name])
        ncomp = 1 if v.ndim == 1 else v.shape[1]
        assert v.shape[0] == n, (name, v.shape, n)
        da = ET.SubElement(cd, "DataArray", type="Float32", Name=name, NumberOfComponents=str(ncomp), format="ascii")
        da.text = _fmt(v)
    da = ET.SubElement(cd, "DataArray", type="Float32", Name="process rank", NumberOfComponents="1", format="ascii")
    da.text = _fmt(np.full(n, float(rank)))
    p = ET.SubElement(ET.SubElement(piece, "Points"), "DataArray", type="Float32", NumberOfComponents="3", format="ascii")
    p.text = _fmt(pts)
    cells = ET.SubElement(piece, "Cells")
    c = ET.SubElement(cells, "DataArray", type="Int32", Name="connectivity", NumberOfComponents="1", format="ascii")
    c.text = "\n          " + "\n          ".join(" ".join(str(int(x)) for x in row) for row in conn) + "\n        "
    o = ET.SubElement(cells, "DataArray", type="Int32", Name="offsets", NumberOfComponents="1", format="ascii")
    o.text = "\n          " + " ".join(str((k + 1) * conn.shape[1]) for k in range(n)) + "\n        "
    t = ET.SubElement(cells, "DataArray", type="UInt8", Name="types", NumberOfComponents="1", format="ascii")
    t.text = "\n          " + " ".join([str(_VTK_TYPE[dim])] * n) + "\n        "
    ET.indent(root, space="  ")
    with open(path, "wb") as f:
        f.write(b'<?xml version="1.0"?>\n')
        ET.ElementTree(root).write(f, encoding="utf-8", xml_declaration=False)
        f.write(b"\n")


def read_vtu(path):
    """Cell data of a (reference- or self-written) ASCII VTU: ordered dict name -> float32 array; plus the number of cells."""
    piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
    n = int(piece.get("NumberOfCells"))
    out = {}
    for da in piece.find("CellData"):
        ncomp = int(da.get("NumberOfComponents", "1"))
        v = np.array(da.text.split(), dtype=np.float32)
        assert v.size == n * ncomp, (da.get("Name"), v.size, n, ncomp)
        out[da.get("Name")] = v.reshape(n, ncomp) if ncomp > 1 else v
    return n, out


def load_solution(path, pv_names):
    """loadSolution (dumux/io/loadsolution.hh:332) for cell-centred models: the primary variables by name, as float64 [n, numEq]."""
    n, data = read_vtu(path)
    missing = [nm for nm in pv_names if nm not in data]
    if missing:
        raise KeyError(f"{path}: no cell data named {missing} (available: {list(data)})")
    return np.stack([data[nm].astype(np.float64) for nm in pv_names], axis=1)


# ------------------------------------------------------------------------------------------------------------------------
# Parallel output / restart (dumux/io/vtkoutputmodule.hh:346 with Dune's parallel VTKWriter; dumux/io/loadsolution.hh:43,332)
# ------------------------------------------------------------------------------------------------------------------------
def piece_name(name, index, nranks, rank):
    """Dune's parallel file names: s<P>-p<r>-<name>-<index>.vtu (the reference's parallel tests compare exactly these files,
    e.g. s0002-p0000-test_richards_lens_tpfa_parallel_yasp-00007.vtu, test/porousmediumflow/richards/lens/CMakeLists.txt:91-112)"""
    return f"s{nranks:04d}-p{rank:04d}-{name}-{index:05d}.vtu"


def pvtu_name(name, index, nranks):
    return f"s{nranks:04d}-{name}-{index:05d}.pvtu"


def write_piece(dirpath, name, index, nranks, rank, node_coords, own_lo, own_hi, fields):
    """This rank's <Piece>: the cells it OWNS (interior partition -- what Dune's VTKWriter iterates over; the rank-0 piece of
    test_richards_lens_tpfa_parallel-reference.vtu holds 12x16 of the 24x16 cells) with their vertices.  `node_coords`: the node
    coordinates of the local box (overlap included), `own_lo/own_hi`: owned index range per axis (local), `fields`: name ->
    array over ALL local cells (x fastest); the overlap cells are cut away here.  Returns the path."""
    import os
    dim = len(node_coords)
    lc = [len(c) - 1 for c in node_coords]
    sl = tuple(slice(int(own_lo[a]), int(own_hi[a])) for a in range(dim - 1, -1, -1))
    owned = {}
    for k, v in fields.items():
        v = np.asarray(v)
        shp = tuple(lc[::-1]) + v.shape[1:]
        w = v.reshape(shp)[sl]
        owned[k] = w.reshape((-1,) + v.shape[1:])
    nodes = [np.asarray(node_coords[a])[int(own_lo[a]):int(own_hi[a]) + 1] for a in range(dim)]
    path = os.path.join(dirpath, piece_name(name, index, nranks, rank))
    write_vtu(path, nodes, owned, rank=rank)
    return path


def write_pvtu(dirpath, name, index, nranks, field_components):
    """The master file rank 0 writes: one <Piece Source=.../> per rank and the declarations of the cell data
    (`field_components`: ordered mapping name -> number of components; `process rank` is appended as in the pieces)."""
    import os
    root = ET.Element("VTKFile", type="PUnstructuredGrid", version="0.1", byte_order="LittleEndian")
    grid = ET.SubElement(root, "PUnstructuredGrid", GhostLevel="0")
    names = list(field_components.keys())
    cd = ET.SubElement(grid, "PCellData", Scalars=names[0] if names else "process rank")
    for nm in names + ["process rank"]:
        ET.SubElement(cd, "PDataArray", type="Float32", Name=nm, NumberOfComponents=str(field_components.get(nm, 1)))
    ET.SubElement(ET.SubElement(grid, "PPoints"), "PDataArray", type="Float32", NumberOfComponents="3")
    for r in range(nranks):
        ET.SubElement(grid, "Piece", Source=piece_name(name, index, nranks, r))
    ET.indent(root, space="  ")
    path = os.path.join(dirpath, pvtu_name(name, index, nranks))
    with open(path, "wb") as f:
        f.write(b'<?xml version="1.0"?>\n')
        ET.ElementTree(root).write(f, encoding="utf-8", xml_declaration=False)
        f.write(b"\n")
    return path


def read_pvtu(path):
    """(list of piece paths, list of cell-data names) of a .pvtu master file"""
    import os
    grid = ET.parse(path).getroot().find("PUnstructuredGrid")
    pieces = [os.path.join(os.path.dirname(path), p.get("Source")) for p in grid.findall("Piece")]
    names = [d.get("Name") for d in grid.find("PCellData")]
    return pieces, names


def load_solution_piece(pvtu_path, rank, pv_names, local_cells, own_lo, own_hi):
    """loadSolution on rank `rank` of a parallel run (dumux/io/loadsolution.hh:332 reads the rank's own piece of the .pvtu; the
    overlap entries are then filled from their owners by the communication of :43 -- here: Engine.halo_exchange after the
    upload).  Returns the LOCAL vector [n_local, numEq] with the owned cells filled and the overlap cells zero."""
    pieces, _ = read_pvtu(pvtu_path)
    own = load_solution(pieces[rank], pv_names)
    dim = len(local_cells)
    out = np.zeros(tuple(int(c) for c in local_cells[::-1]) + (len(pv_names),))
    sl = tuple(slice(int(own_lo[a]), int(own_hi[a])) for a in range(dim - 1, -1, -1))
    out[sl] = own.reshape(out[sl].shape)
    return out.reshape(-1, len(pv_names))
