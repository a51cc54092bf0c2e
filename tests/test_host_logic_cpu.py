"""Host-side logic added in round 2 that needs no GPU: the gather ring's slot life cycle, the operand-format codes, the
bench's workload description and the profile summarisers (run on the committed round-2 artefacts)."""
import importlib.util
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_ring_slot_life_cycle():
    """A slot is handed out in ring order, refused while the loss that read it has not issued its backward, and a stale
    release (the finalizer of an older loss) cannot free a newer gather (mmgclip_b200/distributed.py)."""
    from mmgclip_b200 import distributed as D

    class Handle:
        buffer_ptrs = [1, 2]

    ring = D._GatherRing.__new__(D._GatherRing)
    ring.slots = [D._GatherSlot(torch.empty(1), Handle()) for _ in range(D._GatherRing.RING)]
    ring.i = -1
    got = [ring.acquire() for _ in range(D._GatherRing.RING)]
    assert [g is s for g, s in zip(got, ring.slots)] == [True] * D._GatherRing.RING
    assert ring.acquire() is None                      # every slot still busy: the caller falls back to NCCL
    first = D.GatheredColumns(None, got[0].buf, slot=got[0])
    stale_gen = first.gen
    first.release()
    again = ring.acquire()
    assert again is got[0] and again.busy and again.gen == stale_gen + 1
    got[0].release(stale_gen)                          # finalizer of the OLD loss fires late
    assert again.busy                                  # ... and must not free the new occupant
    again.release(again.gen)
    assert not again.busy


def test_operand_format_codes():
    from mmgclip_b200 import _lib, ops
    a16, b16 = torch.zeros(2, 8, dtype=torch.float16), torch.zeros(2, 8, dtype=torch.float16)
    abf = torch.zeros(2, 8, dtype=torch.bfloat16)
    assert ops._tc_prec("bf16", a16, b16) == _lib.MMG_PREC_F16 == 2
    assert ops._tc_prec("bf16", abf, abf) == _lib.MMG_PREC_BF16 == 1
    assert ops._tc_prec("fp32", a16, abf) == _lib.MMG_PREC_FP32 == 0
    with pytest.raises(ValueError):
        ops._tc_prec("bf16", a16, abf)                 # kind::f16 cannot mix the two 16-bit formats
    with pytest.raises(ValueError):
        ops._tc_prec("bf16", torch.zeros(2, 8), torch.zeros(2, 8))
    assert ops.get_embedding_f16() is False            # library default: bf16 operands


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_launch_summary_reads_the_committed_launch_list():
    ls = _load(os.path.join(ROOT, "scripts", "launch_summary.py"), "launch_summary")
    rows = ls.load(os.path.join(ROOT, "profiles", "r02_ncu_launches_cfg2.csv"))
    ours = [r for r in rows if "mmg::" in r[1]]
    assert len(rows) > 100 and len(ours) > 50
    assert any("infonce_bwd_fused_kernel" in r[1] for r in ours) and any("EpiLseT" in r[1] for r in ours)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", "r02_ncu_launches_n1.csv")], capture_output=True, text=True)
    assert out.returncode == 0 and "infonce_bwd_fused_kernel" in out.stdout and "launches" in out.stdout


def test_committed_sass_summary_names_the_blackwell_instructions():
    text = open(os.path.join(ROOT, "profiles", "sass_summary.md")).read()
    for mnemonic in ("UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "PREEXIT", "ACQBULK"):
        assert mnemonic in text
    line = next(l for l in text.splitlines() if l.startswith("| **total"))
    cells = [c.strip() for c in line.strip("|").split("|")]
    header = [c.strip() for c in next(l for l in text.splitlines() if l.startswith("| kernel")).strip("|").split("|")]
    totals = dict(zip(header, cells))
    assert int(totals["UTCHMMA.2CTA"]) > 0 and int(totals["LDTM"]) > 0 and int(totals["HMMA"]) == 0


def test_bench_describes_the_workload_it_times():
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_mod")
    cfg = bench.workload_config(8, n_sets=4, graph=True, peer=True, symm=True, stored_e=False)
    assert "never" not in cfg["workload"] and "recompute" in cfg["backward"] and "NVLink" in cfg["parallelism"]
    assert "ONE sum" in cfg["small_allreduces"]
    cfg1 = bench.workload_config(1, stored_e=True)
    assert cfg1["backward"].startswith("stored-E")     # an opt-in materialising run is labelled as such
