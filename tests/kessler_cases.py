# -*- coding: utf-8 -*-
"""Shared driver of the K11 fixture (tests/golden/kessler.npz): runs every case through a
backend-specific set of callables and yields (case tag, field name, got, want).  Used by the
oracle test (numpy arrays) and by the GPU parity test (b200 storages)."""
import numpy as np

PREV = ("qc", "qr", "qv", "theta")


def cases(fx, impl, to_dev, to_host, zeros):
    """`impl` provides kessler / saturation_diagnostic / saturation_prognostic / fall_velocity /
    sedimentation / accumulated_precipitation with the oracle's signatures."""
    nx, ny, nz = (int(v) for v in fx["dims"])
    shape = (nx + 1, ny + 1, nz + 1)
    a, k1, k2, dt, sr, rd, rv, cp, lhvw, rhow = (float(v) for v in fx["scalars"])
    beta = rd / rv
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz))
    d = {n: to_dev(fx[n]) for n in ("p_hl", "exn_hl", "p_ml", "exn_ml", "t", "rho", "qv", "qc", "qr",
                                    "h_hl", "vt_in")}
    prev = {n: fx["prev_" + n] for n in PREV}

    for apoil in (True, False):
        pp, ee = (d["p_hl"], d["exn_hl"]) if apoil else (d["p_ml"], d["exn_ml"])
        for evap in (True, False):
            for ow in (True, False):
                o = {n: (zeros(shape) if ow else to_dev(prev[n])) for n in PREV}
                impl.kessler(d["rho"], pp, d["t"], ee, d["qc"], d["qr"], d["qv"], o["qc"], o["qr"],
                             o["qv"], o["theta"], a=a, k1=k1, k2=k2, ow_out_qc_tnd=ow,
                             ow_out_qr_tnd=ow, ow_out_qv_tnd=ow, ow_out_theta_tnd=ow,
                             air_pressure_on_interface_levels=apoil, rain_evaporation=evap,
                             beta=beta, lhvw=lhvw, **box)
                tag = f"kessler_p{int(apoil)}_e{int(evap)}_o{int(ow)}"
                for n in PREV:
                    yield tag, n, to_host(o[n]), fx[f"{tag}_{n}"]
        for ow in (True, False):
            o_qv, o_qc, o_t = zeros(shape), zeros(shape), zeros(shape)
            tnd = zeros(shape) if ow else to_dev(prev["theta"])
            impl.saturation_diagnostic(pp, d["t"], ee, d["qv"], d["qc"], o_qv, o_qc, o_t, tnd, dt=dt,
                                       ow_tnd_theta=ow, air_pressure_on_interface_levels=apoil,
                                       beta=beta, lhvw=lhvw, cp=cp, rv=rv, **box)
            tag = f"satd_p{int(apoil)}_o{int(ow)}"
            for n, v in (("qv", o_qv), ("qc", o_qc), ("t", o_t), ("theta", tnd)):
                yield tag, n, to_host(v), fx[f"{tag}_{n}"]
            o = {n: (zeros(shape) if ow else to_dev(prev[n])) for n in ("qv", "qc", "theta")}
            impl.saturation_prognostic(pp, d["t"], ee, d["qv"], d["qc"], o["qv"], o["qc"], o["theta"],
                                       sr=sr, ow_tnd_qv=ow, ow_tnd_qc=ow, ow_tnd_theta=ow,
                                       air_pressure_on_interface_levels=apoil, beta=beta,
                                       lhvw=lhvw, cp=cp, rv=rv, **box)
            tag = f"satp_p{int(apoil)}_o{int(ow)}"
            for n in o:
                yield tag, n, to_host(o[n]), fx[f"{tag}_{n}"]

    vt = zeros(shape)
    impl.fall_velocity(d["rho"], to_dev(fx["fall_rho_s"]), d["qr"], vt, **box)
    yield "fall_velocity", "vt", to_host(vt), fx["fall_vt"]

    rng = np.random.default_rng(1)
    for order in (1, 2):
        for ow in (True, False):
            # with overwrite the previous content must not matter
            tnd = to_dev(rng.uniform(-1, 1, size=shape) if ow else prev["qr"])
            impl.sedimentation(d["rho"], d["h_hl"], d["qr"], d["vt_in"], tnd, ow_out_tnd_qr=ow,
                               order=order, **box)
            got, want = to_host(tnd), fx[f"sed_{order}_o{int(ow)}"]
            yield f"sed_{order}_o{int(ow)}", "tnd_qr", got[:nx, :ny, :nz], want[:nx, :ny, :nz]

    rho, qr, vt_in = d["rho"], d["qr"], d["vt_in"]
    prec, acc = zeros((nx + 1, ny + 1, 1)), zeros((nx + 1, ny + 1, 1))
    impl.accumulated_precipitation(
        rho[:, :, nz - 1:nz], qr[:, :, nz - 1:nz], vt_in[:, :, nz - 1:nz], to_dev(fx["accprec"])[:, :, :1],
        prec[:, :, :1], acc[:, :, :1], dt=dt, origin=(0, 0, 0), domain=(nx, ny, 1), rhow=rhow)
    yield "precipitation", "prec", to_host(prec), fx["prec"]
    yield "precipitation", "acc", to_host(acc), fx["acc"]
