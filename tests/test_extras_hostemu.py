"""PsiInit / weight masks / Mul iteration through the host-emulated kernel bodies vs the oracle (CPU-only)."""
import pytest

import extras_checks as X


def test_weights_on_device_match_oracle(hostemu_lib, oracle, small_dataset):
    X.check_weights_on_device_match_oracle(hostemu_lib, oracle, small_dataset)


@pytest.mark.parametrize("smooth,osem", [(True, 1.0), (False, 2.5)])
def test_weight_normalisation_variants(hostemu_lib, oracle, small_dataset, smooth, osem):
    X.check_weight_normalisation_variants(hostemu_lib, oracle, small_dataset, smooth, osem)


def test_psi_init_variants(hostemu_lib, oracle, small_dataset):
    X.check_psi_init_variants(hostemu_lib, oracle, small_dataset)


@pytest.mark.parametrize("lam", [0.0, 0.006])
def test_mul_iteration_matches_oracle(hostemu_lib, oracle, small_dataset, lam):
    X.check_mul_iteration_matches_oracle(hostemu_lib, oracle, small_dataset, lam)


def test_affine_blending_weights(hostemu_lib, oracle):
    X.check_affine_blending_weights(hostemu_lib, oracle)
