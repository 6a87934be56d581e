"""CPU tests of host-side logic that needs no device: chunk-copy run grouping, the content key of the per-image device cache,
the NUMA helper's parsing, the bench's voxel sampler and its committed-line schema additions."""
import numpy as np

import bench
from pyradiomics_b200 import featureclasses as FC, numa, voxel


def test_runs_group_consecutive_feature_indices():
    assert voxel._runs([0, 1, 2, 5, 7, 8]) == [(0, 3, 0), (5, 1, 3), (7, 2, 4)]
    assert voxel._runs([]) == [] and voxel._runs([4]) == [(4, 1, 0)]


def test_fingerprint_sees_every_edit_and_ignores_copies():
    rng = np.random.default_rng(0)
    for shape, dt in (((7, 9, 11), np.int16), ((160, 192, 200), np.int16), ((33, 5), np.float32), ((3,), np.uint8)):
        a = rng.integers(0, 200, shape).astype(dt)
        f = FC._fingerprint(a)
        assert FC._fingerprint(a.copy()) == f
        assert FC._fingerprint(np.asfortranarray(a)) == f or a.ndim == 1          # same content, C-ordered by the key
        b = a.copy()
        b.flat[b.size // 3] += 1
        assert FC._fingerprint(b) != f
        c = a.copy()
        c.flat[-1] += 1                                                            # also in the tail bytes / last word
        assert FC._fingerprint(c) != f
    assert FC._fingerprint(np.zeros((4, 4), np.int16)) != FC._fingerprint(np.zeros((2, 8), np.int16))


def test_numa_helper_returns_a_cpu_list_or_nothing():
    cpus = numa.gpu_cpu_affinity(0)
    assert isinstance(cpus, list) and all(isinstance(c, int) for c in cpus)
    assert isinstance(numa.bind_to_gpu(0), list)


def test_bench_voxel_sampler_hits_faces_edges_and_corners():
    v = bench.sample_voxels(64, 4096, 0)
    assert v.shape == (3, 4096) and v.min() == 0 and v.max() == 63
    on_border = ((v == 0) | (v == 63)).sum(0)
    assert (on_border >= 1).mean() > 0.15 and (on_border >= 2).sum() > 50 and (on_border == 3).sum() > 5
    assert np.array_equal(bench.raw_from_levels(np.array([1, 2, 32])), np.array([3, 28, 778], np.int16))


def test_plugin_zchunk_keeps_a_thin_slab_pipelined():
    """the compute / copy pipeline of a class needs several chunks per call: a sixteenth of the planes within [8, 64]
    unless the caller fixes b200_zchunk"""
    from pyradiomics_b200 import featureclasses as FC
    obj = FC.RadiomicsGLCM.__new__(FC.RadiomicsGLCM)
    obj.settings = {}
    assert [obj._zchunk(n) for n in (512, 256, 64, 20, 1)] == [32, 16, 8, 8, 8]
    assert obj._zchunk(4096) == 64
    obj.settings = {"b200_zchunk": 5}
    assert obj._zchunk(512) == 5
