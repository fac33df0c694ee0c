"""phare_b200.h5lite: the h5py surface pyphare's readers and our diagnostics writer use, over a non-HDF5 container"""
import numpy as np
import pytest

from phare_b200 import h5lite


def test_groups_datasets_attributes_round_trip(tmp_path):
    fn = str(tmp_path / "EM_B.h5")
    with h5lite.File(fn, "w") as f:
        f.attrs["cell_width"] = [0.1, 0.2]
        f.attrs["layoutType"] = "yee"
        g = f.require_group("t/0.0000000000/pl0/p0#1")
        g.attrs["lower"] = np.array([0, 4], dtype=np.int32)
        ds = g.create_dataset("EM_B_x", data=np.arange(6.).reshape(2, 3))
        ds.attrs["ghosts"] = 2
        with pytest.raises(ValueError):
            g.create_dataset("EM_B_x", data=np.zeros(1))
    assert h5lite.is_h5lite(fn)
    f = h5lite.File(fn, "r")
    assert f.filename == fn and list(f.keys()) == ["t"] and list(f["t"].keys()) == ["0.0000000000"]
    assert np.array_equal(f.attrs["cell_width"], [0.1, 0.2]) and f.attrs["layoutType"] == "yee"
    p = f["t"]["0.0000000000"]["pl0"]["p0#1"]
    assert p.name == "/t/0.0000000000/pl0/p0#1" and p.name.split("/")[-1] == "p0#1" and "EM_B_x" in p
    assert [k for k, _ in f["/t/0.0000000000/pl0"].items()] == ["p0#1"] and len(f["t/0.0000000000/pl0"].values()) == 1
    d = p["EM_B_x"]
    assert d.shape == (2, 3) and int(d.attrs["ghosts"]) == 2 and np.array_equal(np.asarray(d), np.arange(6.).reshape(2, 3))
    assert np.array_equal(d[:], d[...]) and d[1, 2] == 5.0 and "ghosts" in d.attrs
    with pytest.raises(KeyError):
        f["nope"]
    with pytest.raises(OSError):
        f.create_group("x")
    with pytest.raises(FileNotFoundError):
        h5lite.File(str(tmp_path / "missing.h5"), "r")


def test_appended_records_and_rank_pieces_are_merged(tmp_path):
    fn = str(tmp_path / "ions_charge_density.h5")
    w = h5lite.File(fn, "w")
    w.attrs["dimension"] = 1
    for i in range(3):  # a writer kept open across dumps: one appended record per dump, arrays released after each
        w.require_group(f"t/{i:.10f}/pl0/p0#0").create_dataset("charge_density", data=np.full(4, float(i)))
        w.release_datasets()
        r = h5lite.File(fn)  # a reader in between sees what was flushed so far (and not a stale cached tree)
        assert len(r["t"].keys()) == i + 1
    w.close()
    with h5lite.File(fn + ".rank1", "w") as o:  # what rank 1 wrote
        o.require_group("t/0.0000000000/pl0/p1#1").create_dataset("charge_density", data=np.full(4, 9.))
    f = h5lite.File(fn)
    assert list(f["t"].keys()) == [f"{i:.10f}" for i in range(3)] and int(f.attrs["dimension"]) == 1
    assert list(f["t/0.0000000000/pl0"].keys()) == ["p0#0", "p1#1"]
    assert np.asarray(f["t/2.0000000000/pl0/p0#0/charge_density"])[0] == 2.0
    assert np.asarray(f["t/0.0000000000/pl0/p1#1/charge_density"])[0] == 9.0
    with h5lite.File(fn, "a") as a:  # append mode on an existing file
        a.require_group("t/3.0000000000/pl0/p0#0").create_dataset("charge_density", data=np.zeros(4))
    assert len(h5lite.File(fn)["t"].keys()) == 4


def test_simulator_writes_the_reference_h5_layout(cpu_oracle, tmp_path, monkeypatch):
    """the per-quantity files of diagnostic/detail/h5writer.hpp: names, /t/<time>/pl#/p<rank>#<id>/<dataset>, attributes"""
    import phare_b200.simulator as S
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, two_pop_1d
    monkeypatch.setenv("PHARE_B200_DIAG_FORMAT", "h5")
    monkeypatch.setattr(S, "ops_factory", lambda dim, interp: CpuOps(dim, interp))
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops, bfn, steps=2, largest=[32], diag_dir=str(tmp_path), diag_times=[0.0, 0.005])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    assert sim.dump_diagnostics(0.0, 0.005)
    sim.advance(0.005)
    assert sim.dump_diagnostics(0.005, 0.005)
    sim.close_diagnostics()
    S.dict_instance().stop()
    b = h5lite.File(str(tmp_path / "EM_B.h5"))
    assert list(b["t"].keys()) == ["0.0000000000", "0.0050000000"] and int(b.attrs["interpOrder"]) == 1
    assert list(b.attrs["domain_box"]) == [63] and float(b.attrs["cell_width"][0]) == 0.2
    patches = b["t/0.0050000000/pl0"]
    assert list(patches.keys()) == ["p0#0", "p0#1"]
    p1 = patches["p0#1"]
    assert list(p1.attrs["lower"]) == [32] and list(p1.attrs["upper"]) == [63] and list(p1.keys()) == ["EM_B_x", "EM_B_y", "EM_B_z"]
    assert p1["EM_B_x"].shape == (32 + 1 + 4,) and int(p1["EM_B_x"].attrs["ghosts"]) == 2 and p1["EM_B_y"].shape == (32 + 4,)
    parts = h5lite.File(str(tmp_path / "ions_pop_protons_domain.h5"))["t/0.0050000000/pl0/p0#0"]
    n = parts["weight"].shape[0]
    assert parts["iCell"].shape == (n, 1) and parts["v"].shape == (n, 3) and parts["charge"].shape == (n, 1) and n > 0
    assert float(h5lite.File(str(tmp_path / "ions_pop_beam_flux.h5")).attrs["pop_mass"]) == 2.0


def test_momentum_tensor_of_a_cold_beam(cpu_oracle, tmp_path, monkeypatch):
    """ions/pop/<name>/momentum_tensor (fluid.hpp:97-102: mass * sum w v_i v_j, border-summed): for a population whose
    particles all carry the same velocity, M_ij = mass * n * v_i v_j node by node with n the deposited density; the total
    tensor is the sum over the populations"""
    import pybindlibs.dictator as pp
    import phare_b200.simulator as S
    from oracle.cpu_ops import CpuOps
    from frontend_util import populate, const
    monkeypatch.setenv("PHARE_B200_DIAG_FORMAT", "h5")
    monkeypatch.setattr(S, "ops_factory", lambda dim, interp: CpuOps(dim, interp))
    v0 = (0.5, -0.25, 0.125)
    beam = dict(name="beam", mass=2.0, charge=1.0, ppc=30, seed=5, density=lambda x: 1.0 + 0.3 * np.sin(2 * np.pi * x / 12.8),
                vx=const(v0[0]), vy=const(v0[1]), vz=const(v0[2]), vthx=const(0), vthy=const(0), vthz=const(0))
    core = dict(name="core", mass=1.0, charge=1.0, ppc=30, seed=6, density=const(1.0), vx=const(0), vy=const(0), vz=const(0),
                vthx=const(0.2), vthy=const(0.2), vthz=const(0.2))
    populate([64], [0.2], 2, [beam, core], [const(1.0), const(0.0), const(0.0)], steps=1, largest=[32], diag_dir=str(tmp_path),
             diag_times=[0.0])
    for name, q in (("mt_beam", "/ions/pop/beam/momentum_tensor"), ("mt_core", "/ions/pop/core/momentum_tensor"),
                    ("mt_all", "/ions/momentum_tensor"), ("n_beam", "/ions/pop/beam/density"), ("count", "/particle_count")):
        kind = "info" if name == "count" else "fluid"
        pp.add_string(f"simulation/diagnostics/{kind}/{name}/type", kind)
        pp.add_string(f"simulation/diagnostics/{kind}/{name}/quantity", q)
        pp.add_array_as_vector(f"simulation/diagnostics/{kind}/{name}/write_timestamps", np.array([0.0]))
    sim = S.make_simulator(S.make_hierarchy(), 1, 2, 2)
    sim.initialize()
    assert sim.dump_diagnostics(0.0, 0.005)
    sim.close_diagnostics()
    S.dict_instance().stop()
    at = "t/0.0000000000/pl0"
    Mb = h5lite.File(str(tmp_path / "ions_pop_beam_momentum_tensor.h5"))[at]
    Mc = h5lite.File(str(tmp_path / "ions_pop_core_momentum_tensor.h5"))[at]
    Ma = h5lite.File(str(tmp_path / "ions_momentum_tensor.h5"))[at]
    n = h5lite.File(str(tmp_path / "ions_pop_beam_density.h5"))[at]
    idx = {"x": 0, "y": 1, "z": 2}
    for patch in ("p0#0", "p0#1"):
        dens = np.asarray(n[patch]["density"])
        assert dens[4:-4].min() > 0.5
        for ij in ("xx", "xy", "xz", "yy", "yz", "zz"):
            m = np.asarray(Mb[patch][f"momentum_tensor_{ij}"])
            assert int(Mb[patch][f"momentum_tensor_{ij}"].attrs["ghosts"]) == 4
            want = 2.0 * dens * v0[idx[ij[0]]] * v0[idx[ij[1]]]
            assert np.allclose(m, want, rtol=1e-12, atol=1e-14), ij
            assert np.allclose(np.asarray(Ma[patch][f"momentum_tensor_{ij}"]), m + np.asarray(Mc[patch][f"momentum_tensor_{ij}"]),
                               rtol=1e-13, atol=1e-15)
        assert np.asarray(Mc[patch]["momentum_tensor_xx"])[4:-4].min() > 0  # a warm population has pressure
    cnt = h5lite.File(str(tmp_path / "particle_count.h5"))[at]
    assert int(cnt["p0#0"].attrs["particle_count"]) == 32 * 60 and list(cnt["p0#0"].keys()) == []


def test_tampered_container_is_refused(tmp_path):
    """a container is data, not code: a record that names anything but numpy arrays / plain values is refused"""
    import pickle
    import pytest
    from phare_b200 import h5lite

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned",))
    p = tmp_path / "evil.h5"
    with open(p, "wb") as f:
        f.write(h5lite.MAGIC)
        pickle.dump({"/x": ({}, Evil())}, f)
    with pytest.raises(pickle.UnpicklingError):
        h5lite.File(str(p), "r")
