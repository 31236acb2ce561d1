"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, argument
validation mirrors the reference's error behaviour (no compute calls; no GPU needed), diagnostics."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(cuda_lib):
    from walnuts_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "walnuts_cuda.h")).read()
    declared = set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", hdr)) - {"wn_handle", "wn_config"}
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for s in declared:
        assert hasattr(cuda_lib, s)
    assert cuda_lib.wn_abi_version() == 1
    assert cuda_lib.wn_target_id(b"funnel") == 2 and cuda_lib.wn_target_id(b"nope") < 0


def test_config_struct_layout_matches_header(cuda_lib):
    from walnuts_b200 import _ffi
    assert C.sizeof(_ffi.WnConfig) == 12 * 4 + 6 * 8 + 2 * 8


def test_create_validation_without_gpu(cuda_lib):
    """wn_create validates before touching CUDA: the reference's ValueError cases (walnuts.py:309-320)."""
    from walnuts_b200 import ChainBatch
    for kw in (dict(H0=0.0), dict(M=0), dict(delta=-1.0), dict(minC=3, maxC=2), dict(jitter=1.5)):
        args = dict(target="std_normal", d=3, n_chains=2, H0=0.5, M=5, delta=0.1)
        args.update(kw)
        with pytest.raises(ValueError):
            ChainBatch(**args)


def test_python_callables_are_rejected_loudly(cuda_lib):
    import walnuts_b200 as wb
    with pytest.raises(TypeError, match="no CPU fallback"):
        wb.WALNUTS(lambda q: [0.0, -q], np.zeros(3), numIter=1, warmupIter=0)
    with pytest.raises(TypeError):
        wb.targets.stdGauss(np.zeros(3))
    with pytest.raises(NotImplementedError):
        # orbit statistics of a `generated` that is not a selection of coordinates (index selections are supported)
        wb.WALNUTS(wb.targets.stdGauss, np.zeros(3), generated=lambda q: 2.0 * q[:1], numIter=1, warmupIter=0,
                   recordOrbitStats=True)


def test_user_target_plugin_builds_and_registers(cuda_lib):
    """cuda_target(): nvcc cross-compiles the plug-in here; the C-ABI loads it and hands out a target id; the
    dimension baked into the plug-in is enforced (no compute calls; no GPU needed)."""
    import walnuts_b200 as wb
    from walnuts_b200 import _ffi
    tg = wb.targets.smileDistr.compile()
    assert os.path.isfile(tg.plugin)
    dl = C.CDLL(tg.plugin)
    for sym in ("wn_user_abi", "wn_user_dim", "wn_user_plan", "wn_user_occupancy", "wn_user_launch"):
        assert hasattr(dl, sym)
    assert dl.wn_user_dim() == 2 and dl.wn_user_abi() == cuda_lib.wn_abi_version()
    tid, data = wb.targets.resolve(tg, 2)
    assert tid >= 1000 and data == {} and _ffi.register_user_target(tg.plugin) == tid
    with pytest.raises(ValueError):
        wb.targets.resolve(tg, 3)
    with pytest.raises(ValueError):
        wb.targets.cuda_target("", 513)
    with pytest.raises(RuntimeError, match="nvcc failed"):
        wb.targets.cuda_target("WN_TARGET_LP_GRAD(q, g, data, n_data) { return undefined_symbol; }", 2, name="broken")
    assert cuda_lib.wn_register_user_target(b"/nonexistent.so") < 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from walnuts_b200 import _ffi, build
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_ffi.WalnutsError, match="no CPU fallback"):
        _ffi.load()


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (checked textually)."""
    pkg = os.path.join(ROOT, "walnuts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "/root/reference" not in txt, f


def test_ess_estimator():
    from walnuts_b200 import diagnostics as dg
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 200))
    ess, rhat = dg.ess_bulk(x)
    assert 0.8 * x.size < ess < 1.25 * x.size and abs(rhat - 1) < 0.01
    phi = 0.7
    y = np.zeros((64, 400))
    e = rng.standard_normal((64, 400))
    y[:, 0] = e[:, 0] / np.sqrt(1 - phi ** 2)
    for t in range(1, 400):
        y[:, t] = phi * y[:, t - 1] + e[:, t]
    ess, _ = dg.ess_bulk(y)
    expect = y.size * (1 - phi) / (1 + phi)
    assert 0.85 * expect < ess < 1.15 * expect
    # sufficient statistics are additive across shards (multi-GPU reduction)
    a, b = dg.chain_stats(y[:32], 40), dg.chain_stats(y[32:], 40)
    tot = {k: a[k] + b[k] for k in ("m", "sum_mean", "sum_mean2", "sum_var", "acov_sum")}
    tot["n"] = a["n"]
    whole = dg.ess_from_stats(dg.chain_stats(y, 40))
    assert np.isclose(dg.ess_from_stats(tot)[0], whole[0], rtol=1e-12)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract keys,
    produced by the oracle port on the host cores -- no GPU, no CUDA library involved."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-budget", "1", "--ref-min-transitions", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "grad_evals_per_sec" and line["unit"] == "grad_evals/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("diag_gauss_d1000")
    cb = line["cpu_baseline"]
    # "reference": the real WALNUTS.py is importable here (build container / staged copy); "port": the numpy restatement
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "grad_evals/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_index_generator_detection():
    """recordOrbitStats + generated: only coordinate selections are accepted (api._index_generator); pure host logic."""
    from walnuts_b200.api import _index_generator
    assert list(_index_generator(lambda q: np.array([q[0], q[1]]), 11)) == [0, 1]       # mainFunnel.py:19-20
    assert list(_index_generator(lambda q: q[[3, 0, 3]], 5)) == [3, 0, 3]
    assert list(_index_generator(lambda q: q, 4)) == [0, 1, 2, 3]
    assert _index_generator(lambda q: 2.0 * q, 4) is None
    assert _index_generator(lambda q: np.array([q[0] + q[1]]), 4) is None
    assert _index_generator(lambda q: np.array([np.exp(q[0])]), 4) is None
    assert _index_generator(lambda q: 1 / 0, 4) is None


def test_device_sharding_is_contiguous_with_global_chain_ids():
    """WALNUTS(..., devices=[...]): contiguous blocks of chains, chain_offset = first global chain id of the block."""
    from walnuts_b200.api import _over_devices
    x = np.arange(10 * 3, dtype=np.float64).reshape(10, 3)
    seen = _over_devices(lambda xs, dev, off: (dev, off, xs.copy()), x, [0, 1, 2], 100)
    assert [s[0] for s in seen] == [0, 1, 2]
    assert [s[1] for s in seen] == [100, 103, 106]
    assert np.array_equal(np.concatenate([s[2] for s in seen]), x)
    # fewer chains than devices: one chain per device, the rest unused
    seen = _over_devices(lambda xs, dev, off: (dev, off, xs.shape[0]), x[:2], [0, 1, 2, 3], 0)
    assert seen == [(0, 0, 1), (1, 1, 1)]


def test_bench_workloads_are_the_baseline_configs():
    """bench.py's workload table: shapes of BASELINE.json's five configs (host logic only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    W = bench.WORKLOADS
    assert (W["c2"]["d"], W["c2"]["chains"], W["c2"]["integrator"]) == (1000, 65536, "R2P")
    assert (W["c1"]["d"], W["c1"]["mode"], W["c1"]["H0"], W["c1"]["M"], W["c1"]["delta"]) == (100, "package", 2.0, 10, 0.1)
    assert (W["c3"]["d"], W["c3"]["chains"], W["c3"]["M"]) == (11, 262144, 12)
    assert (W["c4"]["d"], W["c4"]["chains"]) == (100, 16384)
    assert (W["c5"]["d"], W["c5"]["chains"] * 8, W["c5"]["M"], W["c5"]["minC"]) == (756, 1048576, 14, 3)
    assert W["c2_1m"]["chains"] == 1048576
    for name in ("c1", "c3", "c5"):
        q0, data = bench.make_inputs(W[name], 7, 0)
        assert q0.shape == (7, W[name]["d"]) and np.isfinite(q0).all()
    s = bench.sigma_vec()
    assert s.shape == (1000,) and np.isclose(s.min(), 1e-2) and np.isclose(s.max(), 1e2)
    assert np.isclose(s[0], 1e-2) and np.isclose(s[bench.MONITOR - 1], 1e2)      # monitored coordinates span the range


def test_roofline_traffic_comes_from_the_committed_ncu_capture(tmp_path):
    """bench.py never runs under a profiler: roofline.traffic is read from profiles/r02_traffic.json, which
    scripts/traffic_from_ncu.py writes from the CSV of one ncu pass over the bench command."""
    import csv, json, subprocess, sys
    import bench
    t = bench._traffic("c2", 65536, 1)
    assert t["traffic"] and t["traffic"] > 1e9 and "ncu" in t["traffic_source"]
    assert bench._traffic("no_such_workload", 1, 1) == {"traffic": None}
    # the parser on a synthetic log: warm-up + two timed launches per kernel group -> average of launches 2 and 3
    src, dst = tmp_path / "t.csv", tmp_path / "t.json"
    hdr = ["ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size", "Grid Size",
           "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"]
    with open(src, "w", newline="") as fh:
        w = csv.writer(fh, quoting=csv.QUOTE_ALL)
        w.writerow(hdr)
        for i, rd in enumerate((5, 10, 30)):
            for name, val in (("dram__bytes_read.sum", rd), ("dram__bytes_write.sum", 2 * rd), ("gpu__time_duration.sum", 7)):
                w.writerow([i, 1, "python", "h", "void walnutspy_kernel<DiagT, 128, 4, 128, 4, 0, 0, 2, 0, 0>(RunParams)", 1, 7,
                            "(128, 1, 1)", "(592, 1, 1)", 0, "10.0", "s", name, "Mbyte" if "dram" in name else "ns", val])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "scripts", "traffic_from_ncu.py"), str(src), str(dst)], check=True,
                   capture_output=True)
    out = json.load(open(dst))["workloads"]["c2"]
    assert out["dram_read_bytes_per_launch"] == 20e6 and out["traffic_bytes_per_launch"] == 60e6
