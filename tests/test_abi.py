"""CPU: the C-ABI library loads and exports every symbol include/pifpaf_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback exists)."""
import ctypes
import os
import re

import pytest
import torch

from openpifpaf_b200 import _lib, decoder

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'pifpaf_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pifpaf_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f'{name} declared in pifpaf_b200.h but not exported'
    assert set(names) == set(_lib.SYMBOLS), 'ctypes binding and header disagree'


def test_version_and_arch():
    lib = _lib.lib()
    assert lib.pifpaf_abi_version() == 1
    assert lib.pifpaf_build_arch() == b'sm_100a'


def test_default_params_match_reference_statics():
    p = _lib.default_params()
    assert (p.cifhr_neighbors, p.cifhr_threshold, p.seed_threshold, p.caf_score_th) == (16, 0.3, 0.2, 0.3)
    assert (p.keypoint_threshold, p.keypoint_threshold_rel, p.reverse_match, p.greedy) == (0.15, 0.5, 1, 0)
    assert (p.force_complete, p.force_complete_caf_th, p.nms_suppression) == (0, 0.001, 0.00001)
    assert (p.occ_reduction, p.occ_min_scale, p.cifhr_revision) == (2.0, 4.0, 1.0)


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful without a GPU')
def test_no_cpu_fallback():
    d = decoder.CifCaf(17, torch.zeros((19, 2), dtype=torch.int64))
    with pytest.raises(RuntimeError, match='libpifpaf_b200 error'):
        d.call(torch.zeros(17, 5, 3, 3), 16, torch.zeros(19, 8, 3, 3), 16)


def test_bad_arguments_raise_runtime_error():
    with pytest.raises(RuntimeError, match='LongTensor'):
        decoder.CifCaf(17, torch.zeros((19, 2), dtype=torch.int32))
    d = decoder.CifCaf(17, torch.zeros((19, 2), dtype=torch.int64))
    with pytest.raises(RuntimeError, match='float32'):
        d.call(torch.zeros(17, 5, 3, 3, dtype=torch.float64), 16, torch.zeros(19, 8, 3, 3), 16)
    with pytest.raises(RuntimeError, match='initial_ids'):
        d.call_with_initial_annotations(torch.zeros(17, 5, 3, 3), 16, torch.zeros(19, 8, 3, 3), 16,
                                        torch.zeros(1, 17, 4), None)
