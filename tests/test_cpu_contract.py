"""Contract checks that need no GPU: the C header is valid C (and C++), and the reference arm of
bench.py prints exactly one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c_and_cxx(tmp_path):
    hdr = os.path.join(ROOT, "include", "zkb200.h")
    c = tmp_path / "t.c"
    c.write_text('#include "zkb200.h"\nint main(void) { zk_groth16_pkey k; (void)k; return ZK_G1_OUT + ZK_PINOCCHIO_PROOF_OUT > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.dirname(hdr), str(c)])
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "zkb200.h"\nint main() { zk_pinocchio_pkey k{}; return k.n == 0 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(hdr), str(cpp)])


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mpts/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # other ranks of a torchrun job stay silent and exit 0
    env["RANK"] = "1"
    env["WORLD_SIZE"] = "2"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_ocaml_stubs_compile_against_the_c_abi():
    """ocaml/zkb200_stubs.c is shipped as source (no OCaml toolchain here); compile it against the
    real include/zkb200.h with stand-in caml/*.h headers so that a signature drift between the stubs
    and the C ABI is caught on CPU."""
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "tests", "ocaml_mock"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "ocaml", "zkb200_stubs.c")])


def test_bench_cpu_legs_agree():
    """bench.py's two host-side legs on a small prefix: the reference's fold (cpu_baseline) and the
    informational bucket method (cpu_pippenger) give the same point."""
    import bench
    n = 96
    bases_raw, _dl = bench.oracle_bases(n, 7)
    _w, ints = bench.uniform_scalars(n, 11)
    sc_raw = b"".join(v.to_bytes(32, "little") for v in ints)
    _secs, fold = bench.cpu_fold_msm(bases_raw[:64 * 96], sc_raw[:64 * 32], 64, 4)
    line = bench.cpu_pippenger_line(bases_raw, sc_raw, 64, fold, n, 4)
    assert line["agrees_with_fold"] is True and line["value"] > 0 and line["unit"] == "Mpts/s"
    assert line["cores"] == 4 and "not the reference's algorithm" in line["sample"]
    # a sample longer than the buffers is refused, not read out of bounds (bench.py hands slices to C)
    import pytest
    with pytest.raises(ValueError):
        bench.cpu_pippenger_line(bases_raw[:32 * 96], sc_raw[:32 * 32], 64, fold, 32, 4)
    with pytest.raises(ValueError):
        bench.cpu_fold_msm(bases_raw[:32 * 96], sc_raw[:32 * 32], 64, 4)


def test_ocaml_externals_match_the_stubs():
    """Every `external ... = "zkb200_*"` of ocaml/zkb200.ml names a CAMLprim of ocaml/zkb200_stubs.c
    (the .ml file cannot be compiled here), and the README spellings the survey asks for are present."""
    import re
    ml = open(os.path.join(ROOT, "ocaml", "zkb200.ml")).read()
    stubs = open(os.path.join(ROOT, "ocaml", "zkb200_stubs.c")).read()
    prims = set(re.findall(r"CAMLprim value (zkb200_\w+)\(", stubs))
    ext = re.findall(r"external\s+\w+\s*:[^=]+=\s*((?:\"zkb200_\w+\"\s*)+)", ml)
    names = [n for grp in ext for n in re.findall(r"\"(zkb200_\w+)\"", grp)]
    assert len(names) >= 20
    missing = [n for n in names if n not in prims]
    assert not missing, missing
    for needle in ("module Ecp", "module Protocol", "module Test = Test.Make", "Resident.find_or_load keys pk"):
        assert needle in ml, needle
    # no load/free per proof any more ("missing" 4)
    assert "key_free hk" not in ml and "qap_free hq" not in ml


def test_bench_launch_count_helper():
    """bench.py's gpu_launches claim mirrors the tail's chunk-width rule (csrc/msm_impl.cuh:tail)."""
    import bench
    # 2^20 points, c = 17, 20 MSMs per join on 148 SMs: L = 32 -> 2048 chunks -> 32 -> 1
    assert bench.tail_tree_levels(17, 1, 20, 148) == 2
    # the 8-GPU shard: c = 16, L = 32 -> 1024 chunks -> 16 -> 1
    assert bench.tail_tree_levels(16, 1, 20, 148) == 2
    # a single MSM: L = 4 -> 16384 chunks -> 256 -> 4 -> 1
    assert bench.tail_tree_levels(17, 1, 1, 148) == 3
    # tiny window: never below one level
    assert bench.tail_tree_levels(4, 1, 1, 148) == 1
