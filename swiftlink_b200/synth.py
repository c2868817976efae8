"""Synthetic consanguineous pedigree + SNP data in LINKAGE ped/map/dat format.

This produces the workload BASELINE.json names ("synthetic 200-member consanguineous pedigree,
10k SNPs") following SURVEY.md section 8(d): first-cousin and double-first-cousin marriages over
>= 5 generations, about 30 % founders, uniform 0.01 cM marker spacing, MAF ~ U(0.1, 0.5),
genotypes gene-dropped with Haldane recombination from a fixed seed (so they are Mendelian
consistent), bottom two generations typed, and a recessive trait (freq 1e-4) whose affecteds
are autozygous for one founder haplotype at the middle of the map.

The penetrances written to the .dat are 0.001 / 0.001 / 0.999 rather than the examples' fully
penetrant 0 / 0 / 1: with full penetrance the trait likelihood of a sampled descent graph is
exactly zero at positions unlinked to the disease locus (an unaffected relative sharing both
alleles IBD with an affected one), and the reference then stops with "intermediate state had a
likelihood of 0.0 or less" (peeler.cc:92-95) on a pedigree of this size.

File formats are the ones the reference parses (pedigree_parser.cc:80-161, map_parser.cc:11-68,
linkage_parser.cc:18-439).  Everything is deterministic in (seed, n_members, n_markers).
"""
import os

import numpy as np

DEFAULT_SEED = 20261017

MALE, FEMALE = 1, 2


class _Builder(object):
    def __init__(self, rng, target):
        self.rng = rng
        self.target = target
        self.father = []
        self.mother = []
        self.sex = []
        self.gen = []

    @property
    def n(self):
        return len(self.sex)

    def add(self, father, mother, sex, gen):
        self.father.append(father)
        self.mother.append(mother)
        self.sex.append(sex)
        self.gen.append(gen)
        return self.n - 1

    def founder(self, sex, gen):
        return self.add(-1, -1, sex, gen)

    def ancestors(self, i, depth):
        out = set()
        frontier = [i]
        for _ in range(depth):
            nxt = []
            for j in frontier:
                if self.father[j] >= 0:
                    nxt += [self.father[j], self.mother[j]]
            out.update(nxt)
            frontier = nxt
        return out

    def siblings(self, a, b):
        return self.father[a] >= 0 and self.father[a] == self.father[b] and self.mother[a] == self.mother[b]

    def related(self, a, b):
        """share a grandparent or great-grandparent but are not siblings"""
        if self.siblings(a, b) or self.father[a] < 0 or self.father[b] < 0:
            return False
        if {self.father[a], self.mother[a]} & {self.father[b], self.mother[b]}:
            return False            # half siblings
        return len(self.ancestors(a, 3) & self.ancestors(b, 3)) > 0


def _build_structure(rng, target, n_generations, cousin_prob=0.1):
    """Returns (father, mother, sex, gen, n_loops) with exactly `target` members."""
    b = _Builder(rng, target)
    couples = []                    # (father, mother, generation of children)

    # two founding couples whose children intermarry pairwise -> double first cousins below
    fa, ma = b.founder(MALE, 0), b.founder(FEMALE, 0)
    fb, mb = b.founder(MALE, 0), b.founder(FEMALE, 0)
    kids_a = [b.add(fa, ma, s, 1) for s in (MALE, MALE, FEMALE, FEMALE)]
    kids_b = [b.add(fb, mb, s, 1) for s in (FEMALE, FEMALE, MALE, MALE)]
    couples.append((kids_a[0], kids_b[0], 2))
    couples.append((kids_a[1], kids_b[1], 2))
    couples.append((kids_b[2], kids_a[2], 2))
    couples.append((b.founder(MALE, 1), kids_a[3], 2))
    couples.append((kids_b[3], b.founder(FEMALE, 1), 2))
    loops = 0

    g = 2
    while b.n < target:
        last = (g >= n_generations)
        members = []
        for (f, m, gg) in couples:
            if gg != g:
                continue
            remaining = target - b.n
            if remaining <= 0:
                break
            k = int(rng.integers(1, 4)) if g > 2 else int(rng.integers(3, 5))
            k = min(k, remaining)
            for _ in range(k):
                members.append(b.add(f, m, MALE if rng.random() < 0.5 else FEMALE, g))
        if b.n >= target or last:
            if b.n < target and not members:
                # nobody left to have children: extend the bottom generation's couples
                g += 1
                n_generations += 1
            if last and b.n < target:
                # keep adding children to the last generation's couples
                cs = [c for c in couples if c[2] == g]
                while b.n < target:
                    f, m, _ = cs[int(rng.integers(0, len(cs)))]
                    b.add(f, m, MALE if rng.random() < 0.5 else FEMALE, g)
            break

        # marry this generation
        free = list(members)
        rng.shuffle(free)
        married = set()
        if g == 2:
            # force one double-first-cousin marriage (children of the two sibling-pair couples)
            c0 = [x for x in members if b.father[x] == kids_a[0]]
            c1 = [x for x in members if b.father[x] == kids_a[1]]
            pair = [(x, y) for x in c0 for y in c1 if b.sex[x] != b.sex[y]]
            if pair:
                x, y = pair[0]
                married.update([x, y])
                f, m = (x, y) if b.sex[x] == MALE else (y, x)
                couples.append((f, m, g + 1))
                loops += 1
        for a in free:
            if a in married:
                continue
            if rng.random() > 0.85:
                continue                      # stays single (a leaf)
            mate = None
            if rng.random() < cousin_prob:
                cands = [c for c in free if c not in married and c != a and b.sex[c] != b.sex[a]
                         and b.related(a, c)]
                if cands:
                    mate = cands[int(rng.integers(0, len(cands)))]
                    loops += 1
            if mate is None:
                if b.n >= target - 2:
                    continue
                mate = b.founder(FEMALE if b.sex[a] == MALE else MALE, g)
            married.update([a, mate])
            f, m = (a, mate) if b.sex[a] == MALE else (mate, a)
            couples.append((f, m, g + 1))
        g += 1

    # married-in founders whose couple never had a child are disconnected from the pedigree:
    # drop them, then top the bottom generation up to the target size
    father, mother = np.array(b.father), np.array(b.mother)
    sex, gen = np.array(b.sex), np.array(b.gen)
    has_child = np.zeros(b.n, dtype=bool)
    has_child[father[father >= 0]] = True
    has_child[mother[mother >= 0]] = True
    keep = ~((father < 0) & ~has_child)
    remap = np.cumsum(keep) - 1
    father = np.where(father[keep] >= 0, remap[np.maximum(father[keep], 0)], -1)
    mother = np.where(mother[keep] >= 0, remap[np.maximum(mother[keep], 0)], -1)
    sex, gen = sex[keep], gen[keep]
    father, mother, sex, gen = father.tolist(), mother.tolist(), sex.tolist(), gen.tolist()
    bottom = max(gen)
    cs = sorted(set((father[i], mother[i]) for i in range(len(sex)) if gen[i] == bottom and father[i] >= 0))
    while len(sex) < target:
        f, m = cs[int(rng.integers(0, len(cs)))]
        father.append(f); mother.append(m); gen.append(bottom)
        sex.append(MALE if rng.random() < 0.5 else FEMALE)

    return (np.array(father), np.array(mother), np.array(sex), np.array(gen), loops)


def _gene_drop(rng, father, mother, n_markers, theta):
    """Founder-haplotype labels per (person, parental strand, marker); strand 0 = maternal."""
    n = len(father)
    labels = np.full((n, 2, n_markers), -1, dtype=np.int32)
    done = np.zeros(n, dtype=bool)
    for i in range(n):
        if father[i] < 0:
            labels[i, 0, :] = 2 * i
            labels[i, 1, :] = 2 * i + 1
            done[i] = True
    pending = [i for i in range(n) if not done[i]]
    while pending:
        rest = []
        for i in pending:
            if not (done[father[i]] and done[mother[i]]):
                rest.append(i)
                continue
            for strand, parent in ((0, mother[i]), (1, father[i])):
                cross = rng.random(n_markers - 1) < theta
                start = int(rng.integers(0, 2))
                which = (start + np.concatenate(([0], np.cumsum(cross)))) % 2
                labels[i, strand, :] = np.where(which == 0, labels[parent, 0, :], labels[parent, 1, :])
            done[i] = True
        assert len(rest) < len(pending)
        pending = rest
    return labels


# first attempt index that satisfies the constraints, for the default arguments (saves the
# rejection loop; generate() verifies it and falls back to the search if it ever stops holding)
KNOWN_ATTEMPT = {(200, 10000, DEFAULT_SEED): 533}


def generate(n_members=200, n_markers=10000, seed=DEFAULT_SEED, spacing_cm=0.01, n_generations=6,
             trait_freq=1e-4, start_cm=10.0, penetrance=(0.001, 0.001, 0.999),
             loops=(4, 12), min_generation=5, founder_frac=(0.25, 0.36), min_affected=2, cousin_prob=0.1,
             typed_depth=2):
    """Returns a dict describing the pedigree (arrays indexed by file order, ids are 1-based)."""
    hint = KNOWN_ATTEMPT.get((n_members, n_markers, seed))
    attempts = ([hint] if hint is not None else []) + list(range(1000))
    for attempt in attempts:
        rng = np.random.default_rng([seed, attempt])
        father, mother, sex, gen, nloops = _build_structure(rng, n_members, n_generations, cousin_prob)
        n = len(father)
        n_founders = int((father < 0).sum())
        if n != n_members or nloops < loops[0] or nloops > loops[1] or gen.max() < min_generation:
            continue
        if not (founder_frac[0] <= n_founders / float(n) <= founder_frac[1]):
            continue

        theta = 0.5 * (1.0 - np.exp(-2.0 * spacing_cm / 100.0))
        labels = _gene_drop(rng, father, mother, n_markers, theta)

        # trait: one founding haplotype carries the disease allele at the middle marker
        mid = n_markers // 2
        carrier_label = 0               # maternal strand of founder 0
        affected = (labels[:, 0, mid] == carrier_label) & (labels[:, 1, mid] == carrier_label)
        typed = gen >= gen.max() - (typed_depth - 1)          # bottom `typed_depth` generations are genotyped
        if int((affected & typed).sum()) < min_affected:
            continue

        maf = rng.uniform(0.1, 0.5, size=n_markers)
        hap_alleles = (rng.random((2 * n, n_markers)) < maf[None, :]).astype(np.int8) + 1   # 1 or 2
        a_mat = np.take_along_axis(hap_alleles, labels[:, 0, :].astype(np.int64), axis=0) \
            if False else hap_alleles[labels[:, 0, :], np.arange(n_markers)[None, :]]
        a_pat = hap_alleles[labels[:, 1, :], np.arange(n_markers)[None, :]]
        a_mat = np.where(typed[:, None], a_mat, 0).astype(np.int8)
        a_pat = np.where(typed[:, None], a_pat, 0).astype(np.int8)

        pos_cm = start_cm + spacing_cm * np.arange(n_markers)
        return dict(father=father, mother=mother, sex=sex, generation=gen, typed=typed,
                    affected=affected, allele1=a_mat, allele2=a_pat, maf=maf, pos_cm=pos_cm,
                    loops=nloops, n_founders=n_founders, trait_freq=trait_freq, seed=seed,
                    attempt=attempt, penetrance=tuple(penetrance))
    raise RuntimeError("could not build a pedigree meeting the constraints")


def write_linkage(ped, prefix):
    """Writes <prefix>.ped/.map/.dat; returns the three paths."""
    n = len(ped["father"])
    m = len(ped["maf"])
    paths = tuple(prefix + ext for ext in (".ped", ".map", ".dat"))
    d = os.path.dirname(prefix)
    if d and not os.path.isdir(d):
        os.makedirs(d)

    with open(paths[0], "w") as f:
        for i in range(n):
            fa = ped["father"][i] + 1 if ped["father"][i] >= 0 else 0
            mo = ped["mother"][i] + 1 if ped["mother"][i] >= 0 else 0
            aff = 2 if ped["affected"][i] else 1
            g = np.empty(2 * m, dtype=np.int8)
            g[0::2] = ped["allele1"][i]
            g[1::2] = ped["allele2"][i]
            f.write("1 %d %d %d %d %d " % (i + 1, fa, mo, ped["sex"][i], aff))
            f.write(" ".join(map(str, g.tolist())))
            f.write("\n")

    with open(paths[1], "w") as f:
        f.write("#Chr Genpos Marker\n")
        for l in range(m):
            f.write("1 %.4f snp%d\n" % (ped["pos_cm"][l], l + 1))

    with open(paths[2], "w") as f:
        f.write("%d 0 0 5\n" % (m + 1))
        f.write("0 0.0 0.0 0\n")
        f.write(" ".join(str(i + 1) for i in range(m + 1)) + "\n")
        f.write("1 2 # TRAIT\n")
        f.write("%.6f %.6f\n" % (1.0 - ped["trait_freq"], ped["trait_freq"]))
        f.write("1\n")
        f.write("%.4f %.4f %.4f\n" % tuple(ped.get("penetrance", (0.0, 0.0, 1.0))))
        for l in range(m):
            f.write("3 2 # snp%d\n" % (l + 1))
            f.write("%.6f %.6f\n" % (1.0 - ped["maf"][l], ped["maf"][l]))
        f.write("0 0\n")
        f.write(" ".join(["0.1"] + ["0.0001"] * (m - 1)) + "\n")
        f.write("1 0.1 0.45\n")
    return paths


def subset_markers(ped, n_markers):
    """Same pedigree, first n_markers SNPs (used to time the reference where M = 10k does not fit)."""
    out = dict(ped)
    for k in ("allele1", "allele2"):
        out[k] = ped[k][:, :n_markers]
    for k in ("maf", "pos_cm"):
        out[k] = ped[k][:n_markers]
    return out
