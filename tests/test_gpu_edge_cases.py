"""Edge cases of the device path against the oracle: minimum sizes, no genotype data, fully typed
pedigrees, large cutsets (matrices that live in the global slab), explicit prior modes and the
trait-likelihood failure the reference exits on."""
import numpy as np
import pytest

from oracle import orcapi
from swiftlink_b200 import synth, host as H

pytestmark = pytest.mark.gpu


def _setup(tmp_path, ped, iters=20000, X=False):
    from swiftlink_b200 import capi
    paths = synth.write_linkage(ped, str(tmp_path / "case"))
    h = H.Host(*paths, sex_linked=X)
    h.build_peel(iters, seed=5)
    d = h.problem_dict()
    return h, d, orcapi.Problem(d), capi.Plan(d)


def _check_everything(d, orc, plan, sweeps=3, seed=11):
    from swiftlink_b200 import capi
    ch = capi.Chain(plan, seed=seed, chain_id=1)
    ch.lsampler_locus_by_locus(0)
    dg = ch.dg_download()
    ref = np.zeros_like(dg)
    for l in range(orc.M):
        orc.ls_step(ref, l, seed, 1, 0, ignore_left=True, ignore_right=True)
    assert (dg == ref).all()
    for l in sorted(set([0, orc.M // 2, orc.M - 1])):
        r0, m0, p0 = orc.ls_forward(dg, l)
        r1, m1, p1 = ch.debug_forward(l)
        assert r0 == r1 and (m0 == m1).all() and (p0 == p1).all()
    for it in range(1, 1 + sweeps):
        assert orc.ls_sweep(ref, seed, 1, it) == 0
        ch.lsampler_sweep(it)
    dg = ch.dg_download()
    assert (dg == ref).all()
    for itv in sorted(set([0, orc.M - 2])):
        a, b = orc.lod_interval(dg, itv, 0), ch.debug_lod_interval(itv, 0)
        assert (a[0] == b[0]).all() and (a[2] == b[2]).all()
        # prob = ln(result) - recomb - transmission: large terms that nearly cancel when there is no
        # linkage signal, so the tolerance is relative to the terms, not to their difference
        scale = np.abs(np.log(a[0])).max() + abs(orc.marker_transmission())
        assert np.abs(a[1] - b[1]).max() <= 1e-12 * scale
    assert abs(plan.trait_likelihood() - orc.trait_prob()) <= 1e-12 * abs(orc.trait_prob())
    # M-sampler: ordering, whole sweeps and the graph likelihood
    assert (plan.msampler_ordering() == orc.ms_ordering()).all()
    for it in range(100, 100 + sweeps):
        assert orc.ms_sweep(ref, seed, 1, it) == 0
        ch.msampler_sweep(it)
        assert (ch.dg_download() == ref).all()
    want = orc.dg_likelihood(ref)
    assert abs(ch.dg_likelihood() - want) <= 1e-12 * abs(want)
    ch.close()


def _small(**kw):
    args = dict(n_members=24, n_markers=6, seed=3, spacing_cm=1.0, n_generations=4, loops=(0, 6), min_generation=2,
                founder_frac=(0.1, 0.6), min_affected=0, cousin_prob=0.2)
    args.update(kw)
    return synth.generate(**args)


def test_two_markers_minimum(tmp_path):
    ped = synth.subset_markers(_small(), 2)
    h, d, orc, plan = _setup(tmp_path, ped)
    assert orc.M == 2
    _check_everything(d, orc, plan)


def test_no_genotype_data(tmp_path):
    ped = _small()
    ped["allele1"][:] = 0
    ped["allele2"][:] = 0
    h, d, orc, plan = _setup(tmp_path, ped)
    assert (d["elim"] == 15).all()                      # nothing eliminated: every cell of every op is valid
    _check_everything(d, orc, plan)


def test_everyone_typed(tmp_path):
    ped = _small(seed=4, typed_depth=99)
    assert ped["typed"].all()
    h, d, orc, plan = _setup(tmp_path, ped)
    _check_everything(d, orc, plan)


def test_trio(tmp_path):
    ped = dict(father=np.array([-1, -1, 0]), mother=np.array([-1, -1, 1]), sex=np.array([1, 2, 1]),
               affected=np.array([False, False, True]), typed=np.array([True, True, True]),
               allele1=np.array([[1, 2, 1], [1, 2, 2], [1, 2, 1]], np.int8),
               allele2=np.array([[2, 2, 1], [1, 1, 2], [1, 2, 2]], np.int8),
               maf=np.array([0.3, 0.45, 0.2]), pos_cm=np.array([1.0, 2.0, 3.5]), trait_freq=1e-3,
               penetrance=(0.01, 0.01, 0.9))
    h, d, orc, plan = _setup(tmp_path, ped)
    assert orc.N == 3 and orc.F == 2
    _check_everything(d, orc, plan, sweeps=5)


def test_large_cutsets_use_the_global_slab(tmp_path):
    """a heavily inbred pedigree: cutsets of 7+ (16k-cell matrices), arena far beyond shared memory"""
    ped = synth.generate(n_members=70, n_markers=4, seed=21, spacing_cm=1.0, n_generations=5, loops=(14, 40),
                         min_generation=3, founder_frac=(0.05, 0.5), min_affected=0, cousin_prob=0.7)
    h, d, orc, plan = _setup(tmp_path, ped, iters=3000)
    st = plan.stats()
    assert st["max_cutset"] >= 6
    assert st["ls_smem_doubles"] < st["ls_arena_doubles"]            # some matrices live in global memory
    _check_everything(d, orc, plan, sweeps=2)


def test_x_linked_generated(tmp_path):
    ped = _small(seed=9)
    # make the data X-consistent: males homozygous for their maternal allele
    male = ped["sex"] == 1
    ped["allele2"][male] = ped["allele1"][male]
    # fathers pass their single X to daughters: redo daughters' paternal allele
    for i in range(len(ped["sex"])):
        if ped["father"][i] >= 0 and ped["sex"][i] == 2 and ped["typed"][i] and ped["typed"][ped["father"][i]]:
            ped["allele2"][i] = ped["allele1"][ped["father"][i]]
    try:
        h, d, orc, plan = _setup(tmp_path, ped, X=True)
    except RuntimeError:
        pytest.skip("generated genotypes not X-consistent for this seed")
    assert d["sex_linked"] == 1
    _check_everything(d, orc, plan)


def test_strict_mendelian_prior_mode(tmp_path):
    """prior_as_founder given explicitly (founders only): the non-reference, textbook prior"""
    from swiftlink_b200 import capi
    ped = _small(seed=6, n_members=44, n_generations=5, min_generation=4)      # has untyped non-founders
    paths = synth.write_linkage(ped, str(tmp_path / "case"))
    h = H.Host(*paths)
    h.build_peel(20000, seed=5)
    d = h.problem_dict()
    N, M, F = d["N"], d["M"], d["F"]
    strict = np.zeros((N, M, 4))
    for i in range(N):
        for l in range(M):
            strict[i, l] = orcapi.marker_prob(i < F, d["typed"][i], d["genotypes"][i, l], False, d["mapprob"][l])
    assert not np.array_equal(strict, d["marker_prob"])         # the two prior readings differ on this data
    d2 = dict(d)
    d2["marker_prob"] = strict
    d2["prior_as_founder"] = (np.arange(N) < F).astype(np.int32)
    orc, plan = orcapi.Problem(d2), capi.Plan(d2)
    _check_everything(d2, orc, plan)


def test_fully_penetrant_trait_failure_is_reported(tmp_path):
    """peeler.cc:92-95: trait data that are impossible under the model (two affected parents of an
    unaffected child, fully penetrant recessive) make the reference exit with "intermediate state
    had a likelihood of 0.0 or less"; the device reports SLK_ERR_NONPOSITIVE_TRAIT"""
    from swiftlink_b200 import capi
    ped = dict(father=np.array([-1, -1, 0]), mother=np.array([-1, -1, 1]), sex=np.array([1, 2, 1]),
               affected=np.array([True, True, False]), typed=np.array([True, True, True]),
               allele1=np.array([[1, 2, 1], [1, 2, 2], [1, 2, 1]], np.int8),
               allele2=np.array([[2, 2, 1], [1, 1, 2], [1, 2, 2]], np.int8),
               maf=np.array([0.3, 0.45, 0.2]), pos_cm=np.array([1.0, 2.0, 3.5]), trait_freq=1e-3,
               penetrance=(0.0, 0.0, 1.0))
    h, d, orc, plan = _setup(tmp_path, ped)
    with pytest.raises(capi.SlkError) as e:
        plan.trait_likelihood()
    assert e.value.code == capi.ERR_NONPOSITIVE_TRAIT
    ch = capi.Chain(plan, seed=3)
    ch.lsampler_locus_by_locus(0)                      # the marker data themselves are fine
    ch.sync()
    ch.lodscore_accumulate()
    with pytest.raises(capi.SlkError) as e:
        ch.sync()
    assert e.value.code == capi.ERR_NONPOSITIVE_TRAIT
    ch.close()


def _chain_of_families(n_families, n_markers=4, seed=1):
    """a long thin pedigree: family k = founder father, mother (a daughter of family k-1; a founder for k = 0) and
    two children; genotypes gene-dropped without recombination information (each marker independently)"""
    rng = np.random.default_rng(seed)
    father, mother, sex = [-1], [-1], [2]                        # person 0: the first mother
    cur_mother = 0
    for k in range(n_families):
        f = len(father); father.append(-1); mother.append(-1); sex.append(1)
        kids = []
        for c in range(2):
            kids.append(len(father)); father.append(f); mother.append(cur_mother); sex.append(2 if c == 0 else 1)
        cur_mother = kids[0]
    n = len(father)
    father, mother, sex = np.array(father), np.array(mother), np.array(sex)
    maf = rng.uniform(0.2, 0.5, n_markers)
    a1 = np.zeros((n, n_markers), np.int8); a2 = np.zeros((n, n_markers), np.int8)
    for i in range(n):                                           # parents precede children by construction
        for l in range(n_markers):
            if father[i] < 0:
                a1[i, l] = 1 + (rng.random() < maf[l]); a2[i, l] = 1 + (rng.random() < maf[l])
            else:
                a1[i, l] = (a1 if rng.random() < 0.5 else a2)[mother[i], l]
                a2[i, l] = (a1 if rng.random() < 0.5 else a2)[father[i], l]
    typed = father >= 0
    a1[~typed] = 0; a2[~typed] = 0
    return dict(father=father, mother=mother, sex=sex, affected=np.zeros(n, bool), typed=typed, allele1=a1, allele2=a2,
                maf=maf, pos_cm=1.0 + np.arange(n_markers), trait_freq=1e-3, penetrance=(0.01, 0.01, 0.9))


def test_more_than_255_founder_alleles(tmp_path):
    """131 founders: the M-sampler's 16-bit label / 2-word component path (2F > 255)"""
    ped = _chain_of_families(130)
    ped["affected"][[5, 8, 200]] = True
    h, d, orc, plan = _setup(tmp_path, ped, iters=2000)
    assert 2 * orc.F > 255
    _check_everything(d, orc, plan, sweeps=2)


def test_msampler_many_loci_overlapped_launches(tmp_path):
    """4 000 loci (125 likelihood CTAs per hypothesis, every CTA of the chain kernel's cluster busy): the step and
    chain kernels of a sweep are launched with programmatic stream serialisation and the likelihood walk starts
    before the previous pair's chain kernel has finished -- whole M-sweeps still reproduce the oracle's graphs bit
    for bit, interleaved with L-sweeps that invalidate the carried state"""
    from swiftlink_b200 import capi
    ped = synth.generate(n_members=60, n_markers=4000, seed=8, spacing_cm=0.02, n_generations=4, loops=(2, 10),
                         min_generation=2, founder_frac=(0.1, 0.5), min_affected=0, cousin_prob=0.3)
    h, d, orc, plan = _setup(tmp_path, ped, iters=3000)
    ch = capi.Chain(plan, seed=11, chain_id=1)
    ch.lsampler_locus_by_locus(0)
    ref = ch.dg_download()
    for it in (100, 101, 102):
        assert orc.ms_sweep(ref, 11, 1, it) == 0
        ch.msampler_sweep(it)
        assert (ch.dg_download() == ref).all(), it
        if it == 100:
            assert orc.ls_sweep(ref, 11, 1, 7) == 0
            ch.lsampler_sweep(7)
    want = orc.dg_likelihood(ref)
    assert abs(ch.dg_likelihood() - want) <= 1e-12 * abs(want)
    ch.close()
