"""The pooled batching algebra (imsim/photon_pooling.py:227-386), checked the way the
reference's tests/test_photon_pooling.py checks it, against this package's mirror."""
from collections import Counter
from dataclasses import replace
from random import shuffle

import numpy as np

from imsim_b200.photon_array import PhotonArray
from imsim_b200.photon_pooling import LSST_PhotonPoolingImageBuilder as Builder
from imsim_b200.photon_pooling import ObjectInfo, ProcessingMode


def create_fft_obj_list(num_objects, flux=1e6, start_num=0):
    return [ObjectInfo(i + start_num, flux, ProcessingMode.FFT) for i in range(num_objects)]


def create_phot_obj_list(num_objects, flux=1e5, start_num=0):
    return [ObjectInfo(i + start_num, flux, ProcessingMode.PHOT) for i in range(num_objects)]


def create_faint_obj_list(num_objects, flux=100, start_num=0):
    return [ObjectInfo(i + start_num, flux, ProcessingMode.FAINT) for i in range(num_objects)]


def create_mixed_obj_list():
    base_list = create_fft_obj_list(10) + create_phot_obj_list(9) + create_faint_obj_list(1)
    shuffle(base_list)
    return [replace(obj, index=i) for i, obj in enumerate(base_list)]


def _partition_all_same(create_list_fn, desired_mode):
    n_obj, nbatch = 20, 10
    orig = create_list_fn(n_obj)
    fft, phot, faint = Builder.partition_objects(orig, nbatch)
    sizes = {ProcessingMode.FFT: len(fft), ProcessingMode.PHOT: len(phot), ProcessingMode.FAINT: len(faint)}
    for mode, size in sizes.items():
        assert size == (n_obj if mode == desired_mode else 0)
    objects = {ProcessingMode.FFT: fft, ProcessingMode.PHOT: phot, ProcessingMode.FAINT: faint}[desired_mode]
    counts = Counter(o.index for o in objects)
    assert all(counts[o.index] == 1 for o in orig)
    assert all(o.mode == desired_mode for o in objects)


def test_partition_objects_all_fft():
    _partition_all_same(create_fft_obj_list, ProcessingMode.FFT)


def test_partition_objects_all_phot():
    _partition_all_same(create_phot_obj_list, ProcessingMode.PHOT)


def test_partition_objects_all_faint():
    _partition_all_same(create_faint_obj_list, ProcessingMode.FAINT)


def test_partition_objects_mixed_and_phot_below_nbatch():
    objs = create_mixed_obj_list()
    fft, phot, faint = Builder.partition_objects(objs, 10)
    assert (len(fft), len(phot), len(faint)) == (10, 9, 1)
    assert sorted(o.index for o in fft + phot + faint) == list(range(20))
    # PHOT objects with fewer photons than batches are drawn like FAINT ones (photon_pooling.py:374-377)
    few = [ObjectInfo(0, 5, ProcessingMode.PHOT), ObjectInfo(1, 10, ProcessingMode.PHOT)]
    fft, phot, faint = Builder.partition_objects(few, 10)
    assert [o.index for o in phot] == [1] and [o.index for o in faint] == [0]
    assert faint[0].mode == ProcessingMode.PHOT  # the object itself is not modified


def test_make_batches():
    objects = create_fft_obj_list(20)
    for bi, batch in enumerate(Builder.make_batches(objects, 10)):
        assert len(batch) == 2
        assert all(o.index == bi * 2 + i for i, o in enumerate(batch))
    for nbatch, first, big, small in ((6, 2, 4, 3), (9, 2, 3, 2)):
        prev = 0
        batches = list(Builder.make_batches(objects, nbatch))
        assert len(batches) == nbatch
        for bi, batch in enumerate(batches):
            assert len(batch) == (big if bi < first else small)
            assert all(o.index == prev + i for i, o in enumerate(batch))
            prev += len(batch)
        assert prev == 20


def test_make_photon_batches():
    n_phot, n_faint = 15, 5
    phot = create_phot_obj_list(n_phot, start_num=0)
    faint = create_faint_obj_list(n_faint, start_num=n_phot)
    objects = phot + faint
    orig_flux = np.array([o.phot_flux for o in objects])
    nbatch = 11  # does not divide the fluxes
    batches = Builder.make_photon_batches({}, {}, None, phot, faint, nbatch)
    count = Counter(o.index for b in batches for o in b)
    total = np.zeros(len(objects))
    for b in batches:
        for o in b:
            total[o.index] += o.phot_flux
    for i, o in enumerate(objects):
        assert count[i] == (nbatch if o.mode == ProcessingMode.PHOT else 1)
    np.testing.assert_array_almost_equal(total, orig_flux)
    # integer split (f(i+1))//nb - (f i)//nb, photon_pooling.py:302
    f = 123457
    parts = [(f * (i + 1)) // nbatch - (f * i) // nbatch for i in range(nbatch)]
    got = [b[0].phot_flux for b in Builder.make_photon_batches({}, {}, None, [ObjectInfo(0, f, ProcessingMode.PHOT)],
                                                                  [], nbatch)]
    assert got == parts and sum(got) == f
    assert Builder.make_photon_batches({}, {}, None, [], [], 3) == []


def _assert_subbatches(batch, expected_len, subbatches):
    assert len(batch) == sum(len(s) for s in subbatches)
    assert batch == [o for s in subbatches for o in s]
    counts = Counter(o.index for s in subbatches for o in s)
    assert all(counts[o.index] == 1 for o in batch)
    assert [len(s) for s in subbatches] == expected_len


def test_make_photon_subbatches():
    batch = create_phot_obj_list(90) + create_faint_obj_list(10, start_num=90)
    _assert_subbatches(batch, 10 * [10], Builder.make_photon_subbatches(batch, 10))
    _assert_subbatches(batch, 4 * [13] + 4 * [12], Builder.make_photon_subbatches(batch, 8))
    _assert_subbatches(batch, [34] + 2 * [33], Builder.make_photon_subbatches(batch, 3))


def test_merge_photon_arrays():
    class Stamp:
        def __init__(self, pa):
            self.photons = pa

    rng = np.random.default_rng(0)
    stamps = []
    for n in (5, 0, 17, 3):
        pa = PhotonArray(n, x=rng.normal(size=n), y=rng.normal(size=n), flux=np.ones(n),
                         wavelength=rng.uniform(500, 700, n))
        stamps.append(Stamp(pa))
    merged = Builder.merge_photon_arrays(stamps)
    assert merged.size() == 25
    np.testing.assert_array_equal(merged.x, np.concatenate([s.photons.x for s in stamps]))
    np.testing.assert_array_equal(merged.wavelength, np.concatenate([s.photons.wavelength for s in stamps]))
    assert merged.hasAllocatedWavelengths() and not merged.hasAllocatedAngles() and not merged.hasAllocatedPupil()


def test_vectorised_batch_counts_equal_list_algebra():
    from imsim_b200.photon_pooling import photon_batch_counts

    rng = np.random.default_rng(3)
    flux = rng.integers(0, 100000, 300)
    flux[:20] = rng.integers(0, 11, 20)  # fewer photons than batches -> treated as faint
    faint = np.zeros(300, bool)
    faint[50:60] = True
    nbatch = 11
    infos = [ObjectInfo(i, int(f), ProcessingMode.FAINT if faint[i] else ProcessingMode.PHOT)
             for i, f in enumerate(flux)]
    _, phot, fnt = Builder.partition_objects(infos, nbatch)
    draws = list(np.random.default_rng(9).random(len(fnt)))
    it = iter(draws)
    batches = Builder.make_photon_batches({}, {"rng": lambda: next(it)}, None, phot, fnt, nbatch)
    it2 = iter(draws)
    counts = photon_batch_counts(flux, faint, nbatch, lambda: next(it2), clamp=False)
    ref = np.zeros_like(counts)
    for b, batch in enumerate(batches):
        for o in batch:
            ref[b, o.index] += o.phot_flux
    np.testing.assert_array_equal(counts, ref)
    np.testing.assert_array_equal(counts.sum(axis=0), flux)


def test_batching_algebra_matches_the_reference_source():
    """tests/golden/pooling.npz: the reference's own ``partition_objects`` / ``make_photon_batches`` /
    ``make_photon_subbatches`` / ``make_batches`` (their source executed by tests/golden/make_golden_pooling.py on
    recorded uniforms) against the mirror here and against the vectorised ``photon_batch_counts`` that feeds the
    device pipeline: same objects in every batch, in the same order, with the same integer fluxes."""
    import os

    from imsim_b200.photon_pooling import photon_batch_counts

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pooling.npz"))
    mode_of = {0: ProcessingMode.FFT, 1: ProcessingMode.PHOT, 2: ProcessingMode.FAINT}

    def flat(batches):
        off = np.cumsum([0] + [len(b) for b in batches])
        return (off, np.array([o.index for b in batches for o in b], dtype=np.int64),
                np.array([o.phot_flux for b in batches for o in b], dtype=np.float64))

    def check(tag, batches):
        for nm, arr in zip(("off", "idx", "flux"), flat(batches)):
            np.testing.assert_array_equal(arr, g["%s_%s" % (tag, nm)])

    for k in range(int(g["n_cases"])):
        nobj, nbatch = (int(v) for v in g["c%d_in" % k])
        flux, modes, uniforms = g["c%d_flux" % k], g["c%d_modes" % k], g["c%d_uniforms" % k]
        objs = [ObjectInfo(i, int(f), mode_of[int(m)]) for i, (f, m) in enumerate(zip(flux, modes))]
        fft, phot, faint = Builder.partition_objects(objs, nbatch)
        for nm, lst in (("fft", fft), ("phot", phot), ("faint", faint)):
            np.testing.assert_array_equal(np.array([o.index for o in lst], dtype=np.int64), g["c%d_%s" % (k, nm)])
        it = iter(uniforms)
        batches = Builder.make_photon_batches({}, {"rng": lambda: next(it)}, None, phot, faint, nbatch)
        check("c%d_batches" % k, batches)
        check("c%d_fftbatches" % k, list(Builder.make_batches(fft, nbatch)))
        if batches:
            for nsub in (1, 4, 7):
                check("c%d_sub%d" % (k, nsub), Builder.make_photon_subbatches(batches[0], nsub))
        # the vectorised form: counts[b, j] photons of non-FFT object j in batch b
        sel = np.nonzero(modes != 0)[0]
        it = iter(uniforms)
        counts = photon_batch_counts(flux[sel], modes[sel] == 2, nbatch, lambda: next(it), clamp=False)
        want = np.zeros((nbatch, nobj), dtype=np.int64)
        off, idx, fl = g["c%d_batches_off" % k], g["c%d_batches_idx" % k], g["c%d_batches_flux" % k]
        for b in range(len(off) - 1):
            np.add.at(want[b], idx[off[b]:off[b + 1]], fl[off[b]:off[b + 1]].astype(np.int64))
        np.testing.assert_array_equal(counts, want[:, sel])


def test_batch_counts_clamp_like_buildImage_when_bright_objects_are_fewer_than_nbatch():
    """imsim/photon_pooling.py:74,117: the faint partition uses the configured nbatch, the flux split the
    clamped one -- a sparse CCD gets fewer batches (and so fewer boundary recalculations)."""
    from imsim_b200.photon_pooling import photon_batch_counts

    flux = np.array([5, 120000, 7, 33333, 2, 9, 0, 650], dtype=np.int64)  # 3 bright objects for nbatch = 10
    nbatch = 10
    infos = [ObjectInfo(i, int(f), ProcessingMode.PHOT) for i, f in enumerate(flux)]
    _, phot, fnt = Builder.partition_objects(infos, nbatch)
    assert [o.index for o in phot] == [1, 3, 7]
    nb = max(min(nbatch, len(phot)), 1)
    draws = list(np.random.default_rng(2).random(len(fnt)))
    it = iter(draws)
    batches = Builder.make_photon_batches({}, {"rng": lambda: next(it)}, None, phot, fnt, nb)
    it2 = iter(draws)
    counts = photon_batch_counts(flux, np.zeros(len(flux), bool), nbatch, lambda: next(it2))
    assert counts.shape == (3, len(flux))
    ref = np.zeros_like(counts)
    for b, batch in enumerate(batches):
        for o in batch:
            ref[b, o.index] += o.phot_flux
    np.testing.assert_array_equal(counts, ref)
    # no bright object at all: one batch
    c1 = photon_batch_counts(np.array([3, 4]), np.zeros(2, bool), 10, lambda: 0.99)
    assert c1.shape == (1, 2) and c1.sum() == 7
