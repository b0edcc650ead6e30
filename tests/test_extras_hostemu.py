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


@pytest.mark.parametrize("interpolation,with_blending", [(1, True), (1, False), (0, True)])
def test_fuse_group(hostemu_lib, oracle, interpolation, with_blending):
    X.check_fuse_group(hostemu_lib, oracle, interpolation, with_blending)


def test_psf_preparation(hostemu_lib, oracle):
    X.check_psf_preparation(hostemu_lib, oracle)


def test_tiff_io_and_psi_init_from_file(hostemu_lib, oracle, small_dataset, tmp_path):
    X.check_tiff_io_and_psi_init_from_file(hostemu_lib, oracle, small_dataset, tmp_path)


def test_skip_empty_tiles(hostemu_lib, oracle):
    X.check_skip_empty_tiles(hostemu_lib, oracle)


def test_filter_blocks_mirror():
    X.check_filter_blocks_mirror()


def test_n5_io(hostemu_lib, tmp_path):
    X.check_n5_io(hostemu_lib, tmp_path)


def test_debug_interval(hostemu_lib, oracle, small_dataset):
    X.check_debug_interval(hostemu_lib, oracle, small_dataset)


def test_zarr_export(hostemu_lib, tmp_path):
    X.check_zarr_export(hostemu_lib, tmp_path)
