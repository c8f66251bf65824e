"""Converts a B200DISORT_DUMP record file (written by the C++ shim while the REFERENCE's unchanged
examples/amars_sw.cpp ran against this back end on the GPU box: tests/cpp/build/ref_amars_sw) into
tests/golden/amars_sw_real.npz: the real inputs the example builds from data/s8_k_fuller.txt, data/h2so4.txt and
examples/aerosol_output_data.txt (500 wavenumbers x 1 column x 40 layers), the GPU fluxes of that run, and the
CPU oracle's fluxes for the same inputs.      python tests/golden/make_golden_amars.py gpurun_out/amars_sw_dump.bin"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def read_records(path):
    recs, cur = [], None
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            tok = line.decode().split()
            if tok[0] == "forward":
                cur = {"flags": tok[1], "nstr": int(tok[3])}
                recs.append(cur)
                continue
            name, ndim = tok[0], int(tok[1])
            shape = [int(x) for x in tok[2:2 + ndim]]
            n = int(np.prod(shape)) if shape else 1
            cur[name] = np.frombuffer(f.read(8 * n), dtype=np.float64).reshape(shape).copy()
    return recs


def main():
    rec = read_records(sys.argv[1])[0]
    bc = {k[3:]: v for k, v in rec.items() if k.startswith("bc_")}
    flags = O.LAMBER | O.ONLYFL
    ref, _, st = O.forward(rec["prop"], bc, nstr=rec["nstr"], flags=flags, long_double=True)
    assert not st.any()
    err = np.max(np.abs(rec["flx"] - ref) / np.max(np.abs(ref), axis=(2, 3), keepdims=True))
    print("shape", rec["prop"].shape, "flags", rec["flags"], "GPU run vs long-double oracle:", err)
    d = {"prop": rec["prop"], "nstr": rec["nstr"], "flag_tokens": "lamber,quiet,onlyfl", "flx_oracle_ld": ref,
         "flx_gpu_run": rec["flx"], "independent_valid": np.zeros(rec["prop"].shape[:2], dtype=bool)}
    for k, v in bc.items():
        d["bc_" + k] = v
    np.savez_compressed(os.path.join(HERE, "amars_sw_real.npz"), **d)


if __name__ == "__main__":
    main()
