"""The C-ABI boundary without a GPU: the library builds for sm_100a, loads, exports every symbol
include/pm4b200.h declares, its structs have the layout the ctypes mirrors assume, the generated
kernel modules build and carry Blackwell bulk-copy code, and the product refuses to run on CPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from pymc4_b200 import _cabi, _cstructs, codegen, ir as I
from pymc4_b200.ir import lower_model
from tests import models as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pm4b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm4_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = _cabi.build_core()
    lib = _cabi.load_library()
    names = _declared_functions()
    assert len(names) >= 12 and set(names) == set(_cabi.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pm4_abi_version() == I.ABI_VERSION
    dyn = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
    for n in names:
        assert re.search(r"\bT %s\b" % n, dyn), n


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a C probe against the public header and compare sizeof/offsetof with ctypes."""
    structs = {
        "pm4_rv_t": (I.CRV, ["offset", "size", "transform", "shift", "scale"]),
        "pm4_term_t": (I.CTerm, ["kind", "n", "op_begin", "op_end", "n_regs", "value_reg", "param_reg", "plan", "observed",
                                 "ll_offset", "glm_base", "glm_p", "glm_x", "glm_y", "gp_par"]),
        "pm4_op_t": (I.COp, ["opcode", "dst", "a", "b", "mode", "base", "blob", "imm"]),
        "pm4_det_t": (I.CDet, ["n", "op_begin", "op_end", "n_regs", "value_reg", "out_offset"]),
        "pm4_plan_t": (I.CPlan, ["term", "n_gather", "width", "meta", "recs"]),
        "pm4_ir_t": (I.CIR, ["abi_version", "D", "n_rv", "n_terms", "n_ops", "n_dets", "det_row", "rvs", "terms",
                             "ops", "dets", "n_dataf", "dataf", "n_datai", "datai", "struct_hash", "n_plans",
                             "plans", "n_planf", "planf", "n_plani", "plani", "team_warps", "part_reals",
                             "elem_off", "spart_off", "phdr_off"]),
        "pm4_adapt_cfg_t": (_cstructs.AdaptCfg, ["mode", "num_adaptation_steps", "target_accept_prob",
                                                 "exploration_shrinkage", "step_count_smoothing", "decay_rate", "enabled",
                                                 "kind", "adaptation_rate", "mass_windows", "shrinkage_target"]),
        "pm4_hmc_cfg_t": (_cstructs.HMCCfg, ["C", "num_results", "num_burnin", "num_leapfrog_steps", "step_size",
                                             "seed", "adapt", "warps_per_block", "chain_offset", "proposal", "rw_scale",
                                             "step_size_dev"]),
        "pm4_compound_block_t": (_cstructs.CompoundBlock, ["kind", "num_leapfrog_steps", "max_tree_depth", "max_energy_diff",
                                                           "step_size", "rw_scale", "adapt", "dim_scale_dev", "rw_kind_dev"]),
        "pm4_nuts_cfg_t": (_cstructs.NUTSCfg, ["C", "num_results", "num_burnin", "max_tree_depth", "max_energy_diff",
                                               "step_size", "seed", "adapt", "warps_per_block", "chain_offset",
                                               "step_size_dev"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void) {"]
    for cname, (_, fields) in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in fields:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines += ["return 0; }"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout.splitlines())
    for cname, (ct, fields) in structs.items():
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for f in fields:
            assert int(out["%s.%s" % (cname, f)]) == getattr(ct, f).offset, (cname, f)


def test_generated_module_builds_for_sm100a_with_tma_bulk_copy():
    ir, _, _ = lower_model(M.radon_model())
    path = codegen.build_module(ir)
    assert os.path.exists(path) and ir.struct_hash_hex in path and codegen.generator_hash() in path
    lib = ctypes.CDLL(path)
    for sym in ("pm4m_info", "pm4m_logp_grad", "pm4m_dets", "pm4m_init", "pm4m_nuts", "pm4m_hmc",
                "pm4m_da_pooled", "pm4m_export_eps", "pm4m_scratch_bytes", "pm4m_lpt_order"):
        assert hasattr(lib, sym), sym
    elf = subprocess.run(["cuobjdump", "-lelf", path], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    assert "UBLKCP" in sass  # cp.async.bulk (TMA) staging of the plan pools
    assert "SYNCS" in sass  # mbarrier
    assert "ATOMS.CAST" not in sass  # the planned radon likelihood needs no shared-memory CAS loops
    src = codegen.generate_source(ir)
    assert "planned, factored normal" in src and "atomicAdd" not in src


def test_unplanned_gathers_fall_back_to_atomics_in_source():
    idx, y, g = M.poisson_data(n=40)  # below the planning threshold
    ir, _, _ = lower_model(M.poisson_model(idx, y, g))
    assert not ir.plans
    assert "atomicAdd" in codegen.generate_source(ir)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry refuses loudly (BackendUnavailable), it never routes
    through the oracle or any host implementation."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ir, _, _ = lower_model(M.schools_pm4())
    with pytest.raises(_cabi.BackendUnavailable):
        _cabi.DeviceModel(ir)
    import pymc4_b200 as pm

    with pytest.raises(_cabi.BackendUnavailable):
        pm.sample(M.schools_pm4(), num_chains=2, num_samples=2, burn_in=2)
    # the product package never imports, loads or executes anything under oracle/
    pkg = os.path.join(ROOT, "pymc4_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libpm4oracle" not in text, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host cores only (the CPU oracle) and prints ONE JSON line with the
    keys the driver reads; exercised here on a tiny sample so the CPU suite stays fast."""
    import json
    import sys

    env = dict(os.environ, OMP_NUM_THREADS="2")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--burn-in", "5", "--draws", "5", "--cpu-chains", "2"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "leapfrog_grad_evals_per_sec" and d["unit"] == "grad-evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.parametrize("extra", [["--workload", "logit", "--logit-n", "4000"], ["--workload", "gp", "--gp-n", "128"]])
def test_bench_reference_arm_of_the_dense_workloads(extra):
    """The CPU arm of the LOGIT and GP workloads (reduced sizes here): same contract line, bounded samples."""
    import json
    import sys

    env = dict(os.environ, OMP_NUM_THREADS="2")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--cpu-chains", "2"] + extra,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "leapfrog_grad_evals_per_sec" and d["value"] > 0
    assert d["dtype"] == ("f64" if "gp" in extra else "f32")
    assert d["cpu_baseline"]["kind"] == "port" and d["config"]["num_leapfrog_steps"] == 3
