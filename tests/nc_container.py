"""TEST INFRASTRUCTURE: writer of the flat container the NetCDF stand-in reads (tests/cpp/shim_nc/netcdf.h), used to
give the reference's RFM attenuator and read_weights_rfm a synthetic `amarsw-ck-B1.nc` (the real file is absent
from the reference tree)."""
import struct

import numpy as np


def write(path, dims, variables):
    """dims: {name: len}; variables: {name: ndarray} (stored as float64, C order)."""
    head = b"B200NC01" + struct.pack("<i", len(dims))
    for k, n in dims.items():
        head += k.encode().ljust(64, b"\0") + struct.pack("<q", int(n))
    head += struct.pack("<i", len(variables))
    off = len(head) + len(variables) * (64 + 16)
    table, payload = b"", b""
    for k, a in variables.items():
        a = np.ascontiguousarray(a, dtype="<f8")
        table += k.encode().ljust(64, b"\0") + struct.pack("<qq", a.size, off + len(payload))
        payload += a.tobytes()
    with open(path, "wb") as f:
        f.write(head + table + payload)


def write_ck_file(path, bw):
    """amarsw-ck-B1.nc-shaped container from a BandWorkload (pyharp_b200.workloads.make_band): dims / variables as
    src/opacity/rfm.cpp:40-112 and src/utils/read_weights.cpp:26-38 read them."""
    t0 = bw.tables[0]
    nw, npres, ntemp = t0["kdata"].shape
    kaxis = t0["kaxis"]
    dims = {"Wavenumber": nw, "Pressure": npres, "TempGrid": ntemp, "weights": nw}
    var = {"Wavenumber": kaxis[:nw], "Pressure": np.exp(kaxis[nw:nw + npres]), "TempGrid": kaxis[nw + npres:],
           "Temperature": t0["krefatm"][1], "weights": bw.weights}
    for name, t in zip(("CO2", "H2O"), bw.tables):
        var[name] = t["kdata"]
    write(path, dims, var)
