"""CPU: the chunk records the fused tensor-core kernel reads from its parameters (host-built, csrc/mlp_tc.cu:
tc_chunk_records) -- one record per weight-ring slot.  The MMA issuer, the weight producer and the relay lane of the kernel
walk exactly these records, so their invariants are the kernel's pipeline protocol: every byte of a step's weight image is
streamed once and in order, every readiness barrier a step depends on is waited for exactly once, an 8-bit remainder chunk
always follows the 16-bit chunk of the same columns, and a step fits the table.  Host code only: no GPU, no launch."""
import ctypes as C

import numpy as np
import pytest

import vfn_testutil as U
from vfnerf_b200 import _lib

MAX_CHUNKS, TABLE_STEPS = 20, 13
BAR_AUX, BAR_SKIP, BAR_EMB0, BAR_AUX_STATIC = 4, 5, 6, 7
PROGRAMS = {"render": 0, "vf_full": 1, "v_only": 2, "dgrad": 3, "dgrad_vf": 4}


@pytest.fixture(scope="module")
def dbg():
    _lib.build_debug()
    return _lib.debug_lib()


def _table(dbg, precision, program, keep=0):
    case, z = U.load_golden("full_det")
    model = U.make_model(case, U.case_state(case, z), "cpu", precision=precision)
    cfg = model._render_cfg(1024, False)
    va, ra = model.vector_field_network.arena(), model.rendering_network.arena()
    rec = np.zeros((TABLE_STEPS, MAX_CHUNKS, 4), np.uint32)
    nch = np.zeros(TABLE_STEPS, np.int32)
    facts = np.zeros((TABLE_STEPS, 6), np.int32)
    ns = np.zeros(1, np.int32)
    _lib.check_debug(dbg.vfnerf_debug_chunk_table(C.byref(cfg), C.byref(va.desc), C.byref(ra.desc), keep, PROGRAMS[program],
                                                  rec.ctypes.data, nch.ctypes.data, facts.ctypes.data, ns.ctypes.data),
                     "vfnerf_debug_chunk_table")
    n = int(ns[0])
    return rec[:n], nch[:n], facts[:n]


CASES = [("bf16", "render", 0), ("bf16", "vf_full", 0), ("bf16", "v_only", 0), ("bf16", "render", 1), ("bf16", "dgrad", 1),
         ("bf16x3", "render", 0), ("bf16x3", "vf_full", 0), ("bf16x3", "v_only", 0),
         ("fp16f8", "render", 0), ("fp16f8", "vf_full", 0), ("fp16f8", "v_only", 0)]


@pytest.mark.parametrize("precision,program,keep", CASES)
def test_chunk_records_cover_every_step_exactly_once(dbg, precision, program, keep):
    rec, nch, facts = _table(dbg, precision, program, keep)
    n_steps = {"render": 13, "vf_full": 9, "v_only": 8, "dgrad": 13}[program]
    assert len(nch) == n_steps <= TABLE_STEPS
    slot = 32768 if precision == "bf16" else 16384
    for si in range(n_steps):
        N, K, chunk_k, n_seg, fresh, use_lo = (int(v) for v in facts[si])
        nc = int(nch[si])
        assert 1 <= nc <= MAX_CHUNKS
        r = rec[si, :nc].astype(np.int64)
        a_off, kc, need = r[:, 0], r[:, 1] & 0xFFFF, r[:, 1] >> 16
        nbytes, off16, flags = r[:, 2] & 0xFFFF, r[:, 2] >> 16, r[:, 3]
        # one "last" flag, on the last record
        assert (flags[:-1] & 1).sum() == 0 and flags[-1] & 1
        # a record is this CTA's half of a K chunk: N/2 rows x kc columns x 2 bytes, never more than a ring slot
        assert (nbytes == (N // 2) * kc * 2).all() and (nbytes <= slot).all() and (kc % 16 == 0).all() and (kc <= chunk_k).all()
        # the producer streams the half image front to back; images of steps whose lo products are not issued (the feature
        # step inside render()) keep the W_lo chunks in between, so offsets may skip but never go back or overlap
        assert off16[0] == 0 and (np.diff(off16) >= nbytes[:-1] // 16).all()
        assert off16[-1] * 16 + nbytes[-1] <= (N // 2) * K * 2
        if use_lo or precision == "bf16":
            assert (np.diff(off16) == nbytes[:-1] // 16).all() and off16[-1] * 16 + nbytes[-1] == (N // 2) * K * 2
        # every readiness barrier the step depends on is waited for exactly once, by the first chunk that reads its columns
        seen = 0
        for m in need:
            assert int(m) & seen == 0 and int(m) & ~fresh == 0 and bin(int(m)).count("1") <= 2
            seen |= int(m)
        assert seen == fresh
        # A offsets are K-slab units of the 128-row tile (128 x 16 bytes = 128 sixteen-byte units per slab)
        assert (a_off % 128 == 0).all()
        if precision == "fp16f8":
            # an 8-bit remainder chunk (flag bit 1) directly follows the 16-bit chunk of the same K columns
            for c in np.nonzero(flags & 2)[0]:
                assert c > 0 and not flags[c - 1] & 2 and kc[c] == kc[c - 1] and need[c] == 0 and kc[c] % 32 == 0
        else:
            assert (flags & 2).sum() == 0
        if precision == "bf16":
            assert (flags >> 8).sum() == 0          # no hi / lo copies in the plain tile


def test_split_precision_render_program_shape(dbg):
    """The shipped nets in the two split-precision modes: 8 hidden VF steps, the feature step, 4 colour steps; the colour net's
    first step waits for the aux barriers through its explicit segment mask (its columns alias lo columns), and a hidden VF
    layer is 4 x (hi + lo chunk) + the bias chunk = 9 ring slots."""
    for precision in ("bf16x3", "fp16f8"):
        rec, nch, facts = _table(dbg, precision, "render")
        assert [int(f[0]) for f in facts] == [256, 256, 256, 224, 256, 256, 256, 256, 256, 256, 256, 256, 256]
        assert int(nch[1]) == int(nch[2]) == 9
        aux_bits = (1 << BAR_AUX) | (1 << BAR_AUX_STATIC)
        need9 = [int(v) >> 16 for v in rec[9, :int(nch[9]), 1]]
        assert sum(1 for m in need9 if m == aux_bits) == 1 and need9[0] == 1
        assert int(facts[9][4]) == 0xF | aux_bits
        # layer 0 and the skip layer read the (hi | lo) embedding: barrier 6, only in step 0 (the skip layer reuses the phase)
        assert (int(rec[0, 0, 1]) >> 16) == 1 << BAR_EMB0
        assert all(((int(v) >> 16) >> BAR_EMB0) & 1 == 0 for v in rec[4, :int(nch[4]), 1])
        # dual chunks (W_hi against the hi AND the lo copy of A): every hidden VF step in bf16x3; in fp16f8 only the two steps
        # that read the small (hi | lo) embedding segment, which keeps three 16-bit products
        dual_steps = [si for si in range(13) if (rec[si, :int(nch[si]), 3] >> 8 != 0).any()]
        assert dual_steps == (list(range(8)) if precision == "bf16x3" else [0, 4])
