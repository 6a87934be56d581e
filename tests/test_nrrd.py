"""Output assembly: the NRRD writer against the reader that parses the reference's own
bundled test data (oracle/ref_harness.read_nrrd)."""
import gzip
import os

import numpy as np
import pytest

import ref_harness as rh
from pyradiomics_b200 import nrrd


@pytest.mark.parametrize("compress", [True, False])
def test_roundtrip_through_the_reference_data_reader(tmp_path, compress):
    rng = np.random.default_rng(4)
    a = rng.normal(size=(5, 6, 7))
    a[0, 0, 0] = np.nan
    p = nrrd.write_nrrd(str(tmp_path / "m.nrrd"), a, spacing_xyz=(0.78125, 0.78125, 6.5), compress=compress)
    b, sp = rh.read_nrrd(p)
    assert b.dtype == np.float64 and b.shape == a.shape
    assert np.array_equal(a, b, equal_nan=True)
    assert sp == pytest.approx((0.78125, 0.78125, 6.5))
    head = open(p, "rb").read().split(b"\n\n", 1)[0].decode()
    assert head.startswith("NRRD0004") and "sizes: 7 6 5" in head and "type: double" in head
    assert ("encoding: gzip" in head) == compress


def test_gzip_member_is_standard(tmp_path):
    a = np.arange(2 * 3 * 4, dtype=np.int16).reshape(2, 3, 4)
    p = nrrd.write_nrrd(str(tmp_path / "i.nrrd"), a, chunk_bytes=16)         # several deflate chunks
    body = open(p, "rb").read().split(b"\n\n", 1)[1]
    assert np.array_equal(np.frombuffer(gzip.decompress(body), dtype="<i2").reshape(a.shape), a)


def test_write_maps_names_like_the_reference_keys(tmp_path):
    maps = {"ngtdm": np.random.default_rng(0).normal(size=(5, 3, 4, 5)), "gldm": np.zeros((2, 3, 4, 5))}
    names = {"ngtdm": ["Busyness", "Coarseness", "Complexity", "Contrast", "Strength"], "gldm": ["A", "B"]}
    out = nrrd.write_maps(str(tmp_path / "o"), maps, names, spacing_xyz=(1, 2, 3), workers=3)
    assert set(out) == {f"original_ngtdm_{n}" for n in names["ngtdm"]} | {"original_gldm_A", "original_gldm_B"}
    b, sp = rh.read_nrrd(out["original_ngtdm_Contrast"])
    assert np.array_equal(b, maps["ngtdm"][3]) and sp == pytest.approx((1, 2, 3))
    assert os.path.basename(out["original_gldm_B"]) == "original_gldm_B.nrrd"
