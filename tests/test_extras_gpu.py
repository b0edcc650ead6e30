"""PsiInit / weight masks / Mul iteration on the GPU through the C ABI vs the oracle."""
import pytest

import extras_checks as X

pytestmark = pytest.mark.gpu


def test_weights_on_device_match_oracle(product_lib, oracle, small_dataset):
    X.check_weights_on_device_match_oracle(product_lib, oracle, small_dataset)


@pytest.mark.parametrize("smooth,osem", [(True, 1.0), (False, 2.5)])
def test_weight_normalisation_variants(product_lib, oracle, small_dataset, smooth, osem):
    X.check_weight_normalisation_variants(product_lib, oracle, small_dataset, smooth, osem)


def test_psi_init_variants(product_lib, oracle, small_dataset):
    X.check_psi_init_variants(product_lib, oracle, small_dataset)


@pytest.mark.parametrize("lam", [0.0, 0.006])
def test_mul_iteration_matches_oracle(product_lib, oracle, small_dataset, lam):
    X.check_mul_iteration_matches_oracle(product_lib, oracle, small_dataset, lam)


def test_affine_blending_weights(product_lib, oracle):
    X.check_affine_blending_weights(product_lib, oracle)


@pytest.mark.parametrize("interpolation,with_blending", [(1, True), (1, False), (0, True)])
def test_fuse_group(product_lib, oracle, interpolation, with_blending):
    X.check_fuse_group(product_lib, oracle, interpolation, with_blending)


def test_psf_preparation(product_lib, oracle):
    X.check_psf_preparation(product_lib, oracle)


def test_skip_empty_tiles(product_lib, oracle):
    X.check_skip_empty_tiles(product_lib, oracle)


def test_tiff_io_and_psi_init_from_file(product_lib, oracle, small_dataset, tmp_path):
    X.check_tiff_io_and_psi_init_from_file(product_lib, oracle, small_dataset, tmp_path)


def test_n5_io(product_lib, tmp_path):
    X.check_n5_io(product_lib, tmp_path)


def test_debug_interval(product_lib, oracle, small_dataset):
    X.check_debug_interval(product_lib, oracle, small_dataset)


def test_zarr_export(product_lib, tmp_path):
    X.check_zarr_export(product_lib, tmp_path)
