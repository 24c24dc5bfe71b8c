"""HBondAnalysis -- mirror of the reference's hydrogen-bond analysis (src/system/hbonds.rs) on the GPU path.

Selections, bonded-atom lookups and the bookkeeping of chains and pairs stay on the host, as in the reference (HBondChain,
HBondChainGroups::new hbonds.rs:108-150, sanity_check_pairs :337-370); the per-frame search -- acceptors binned into a cell
grid, donors walking their neighbourhood, the distance and the donor-hydrogen-acceptor angle criterion (:160-335) -- runs in
one kernel per (acceptor grid, donor list) and frame batch (groan_gpu_hbonds)."""
from collections import OrderedDict


class HBondError(Exception):
    pass


class HBondChain:
    """acceptors / donors / hydrogens of one chain, as groups of the System (the reference takes selection queries; the
    selection language stays outside this package)"""

    def __init__(self, acceptors, donors, hydrogens):
        self.acceptors, self.donors, self.hydrogens = acceptors, donors, hydrogens


class HBondAnalysis:
    """HBondAnalysis::new(system, chains, pairs, max_distance, min_angle) (hbonds.rs:372-420) and FrameAnalyze::analyze (:160-210).
    `bonds`: iterable of (i, j) bonded pairs used to find the hydrogens of every donor (System::bonded_atoms_iter); defaults to
    the bonds the System already holds (add_bonds / guess_bonds)."""

    def __init__(self, system, chains, pairs, max_distance, min_angle, bonds=None):
        if not chains:
            raise HBondError("NoChains")
        if bonds is None:
            bonds = getattr(system, "_bonds", None)
        if bonds is None:
            raise HBondError("NoBonds: HBondAnalysis needs the topology (System.add_bonds or System.guess_bonds)")
        bonded = {}
        for a, b in bonds:
            bonded.setdefault(int(a), []).append(int(b))
            bonded.setdefault(int(b), []).append(int(a))
        self.system = system
        self.max_distance, self.min_angle = float(max_distance), float(min_angle)
        self.chains = []
        for ch in chains:
            acc = system._group(ch.acceptors)
            hyd = set(int(i) for i in system._group(ch.hydrogens).index_array(system.n_atoms))
            donors = []
            for d in system._group(ch.donors).index_array(system.n_atoms):
                hs = [h for h in sorted(bonded.get(int(d), [])) if h in hyd]
                if hs:  # donors without a bonded hydrogen are ignored (hbonds.rs:130-133)
                    donors.append((int(d), hs))
            if len(acc) == 0 and not donors:
                raise HBondError("EmptyChain")
            self.chains.append((ch.acceptors, donors))
        seen, used = set(), set()
        for c1, c2 in pairs:  # sanity_check_pairs (hbonds.rs:337-370)
            for c in (c1, c2):
                if c >= len(self.chains):
                    raise HBondError("InvalidPair(%d, %d)" % (c1, c2))
                used.add(c)
            key = (min(c1, c2), max(c1, c2))
            if key in seen:
                raise HBondError("DuplicatePair(%d, %d)" % (c1, c2))
            seen.add(key)
        self.pairs = [(int(a), int(b)) for a, b in pairs]

    def analyze(self):
        """one HBondMap per frame of the current batch: OrderedDict {(chain1, chain2): records}"""
        s, F = self.system, self.system.n_frames
        maps = [OrderedDict() for _ in range(F)]
        import numpy as np
        for c1, c2 in self.pairs:
            if c1 == c2:
                parts = [s.hbonds_single(self.chains[c1][0], self.chains[c1][1], self.max_distance, self.min_angle)]
            else:  # analyze_pair (hbonds.rs:214-238): acceptors of chain 1 with donors of chain 2, then the other way round
                parts = [s.hbonds_single(self.chains[c1][0], self.chains[c2][1], self.max_distance, self.min_angle),
                         s.hbonds_single(self.chains[c2][0], self.chains[c1][1], self.max_distance, self.min_angle)]
            for f in range(F):
                maps[f][(c1, c2)] = np.concatenate([p[f] for p in parts])
        return maps
