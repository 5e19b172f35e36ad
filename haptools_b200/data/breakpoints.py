"""
In-memory model of a .bp file (haptools/data/breakpoints.py:14-60) and the two methods the
ancestry transform needs from it:

    Breakpoints.encode            haptools/data/breakpoints.py:162-200
    Breakpoints.population_array  haptools/data/breakpoints.py:259-330   (device kernel here)

`data` maps sample -> [blocks of strand 1, blocks of strand 2]; a block array has the fields
pop, chrom, bp (end position), cm.  .bp parsing/writing is out of scope.
"""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np

HapBlock = [("pop", "U6"), ("chrom", "U10"), ("bp", np.uint32), ("cm", np.float64)]
HapBlockEncoded = [("pop", np.uint8), ("chrom", "U10"), ("bp", np.uint32), ("cm", np.float64)]


class Breakpoints:
    def __init__(self, fname: Path | str = None, log: logging.Logger = None):
        self.fname = Path(fname) if isinstance(fname, str) else fname
        self.data = None
        self.labels = None
        self.log = log or logging.getLogger(f"haptools_b200.{type(self).__name__}")

    def encode(self, labels: tuple = None):
        """Replace population strings by integers, in place; first-seen order unless `labels`
        fixes it (haptools/data/breakpoints.py:162-200)."""
        if self.labels is not None:
            raise ValueError("The data has already been encoded.")
        table = {} if labels is None else {pop: i for i, pop in enumerate(labels)}
        seen = set()
        for blocks in self.data.values():
            for strand in range(len(blocks)):
                old = blocks[strand]
                new = np.zeros(len(old), dtype=HapBlockEncoded)
                for name in ("chrom", "bp", "cm"):
                    new[name] = old[name]
                for i, pop in enumerate(old["pop"]):
                    pop = str(pop)
                    if pop not in table:
                        table[pop] = len(table)
                    new["pop"][i] = table[pop]
                    seen.add(pop)
                blocks[strand] = new
        self.labels = {k: v for k, v in table.items() if k in seen}

    # ------------------------------------------------------------------ device path
    def _flatten(self, variants: np.ndarray, samples):
        data = self.data if samples is None else {s: self.data[s] for s in samples}
        names = tuple(data.keys())
        chroms = {}
        for c in variants["chrom"]:
            chroms.setdefault(str(c), len(chroms))
        var_chrom = np.fromiter((chroms[str(c)] for c in variants["chrom"]), dtype=np.uint64, count=len(variants))
        var_key = (var_chrom << np.uint64(32)) | variants["pos"].astype(np.uint64)
        keys, pops, seg_off = [], [], [0]
        for blocks in data.values():
            for strand in blocks:
                # blocks on chromosomes the genotypes do not use can never be hit: drop them;
                # group the rest by chromosome, keeping the file order within a chromosome
                # (the reference selects `blocks[blocks["chrom"] == chrom]`)
                cidx = np.fromiter((chroms.get(str(c), -1) for c in strand["chrom"]), dtype=np.int64, count=len(strand))
                keep = np.nonzero(cidx >= 0)[0]
                order = keep[np.argsort(cidx[keep], kind="stable")]
                keys.append((cidx[order].astype(np.uint64) << np.uint64(32)) | strand["bp"][order].astype(np.uint64))
                pops.append(strand["pop"][order])
                seg_off.append(seg_off[-1] + len(order))
        keys = np.concatenate(keys) if keys else np.zeros(0, dtype=np.uint64)
        pops = np.concatenate(pops) if pops else np.zeros(0, dtype=np.uint8)
        return names, list(chroms), var_key, keys, pops.astype(np.uint8), np.asarray(seg_off, dtype=np.uint32)

    def population_provider(self, variants: np.ndarray, samples: tuple = None, device=None):
        """The device side of `population_array`, column chunk by column chunk: returns an object whose
        `labels(cols)` launches ht_bp_ancestry for the variants `variants[cols]` and returns their labels as a CUDA
        uint8 tensor `(n, len(cols), 2)` (asynchronous on the current stream), and whose `check()` raises the
        reference's ValueError if some (sample, strand, variant) had no block (haptools/data/breakpoints.py:
        296-323).  The block tables are uploaded once.  This is what the streamed PGEN path asks for chunk by
        chunk (stream.transform_pgen_stream), so the `(n, p, 2)` ancestry matrix the reference materialises
        (haptools/transform.py:658) never exists."""
        import torch

        from .. import _cuda, engine

        engine.require_cuda()
        if self.labels is None:
            raise ValueError("population_tensor needs encoded breakpoints: call encode() first")
        names, chroms, var_key, keys, pops, seg_off = self._flatten(variants, samples)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n, p = len(names), len(variants)
        self.log.info(f"Obtaining ancestry for {n} samples and {p} variants in {len(chroms)} chromosomes")
        owner = self

        class Provider:
            n_samples = n

            def __init__(self):
                with torch.cuda.device(dev):
                    up = lambda a, t: torch.from_numpy(np.ascontiguousarray(a).view(t)).to(dev)  # noqa: E731
                    self.d_keys, self.d_pops = up(keys, np.int64), up(pops, np.uint8)
                    self.d_seg, self.d_var = up(seg_off, np.int32), up(var_key, np.int64)
                    self.bad = torch.full((1,), -1, dtype=torch.int64, device=dev)
                    # chunk-local failure word: (strand index << 32 | chunk column); translated back on the host
                    self.pending = []

            def labels(self, cols=None, variant_major: bool = False):
                with torch.cuda.device(dev):
                    if cols is None:
                        d_var, c = self.d_var, p
                    else:
                        idx = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int64)).to(dev)
                        d_var, c = self.d_var[idx], len(cols)
                    if variant_major:  # (c, n, 2) memory, seen as (n, c, 2): the layout of a pgenlib chunk
                        anc = torch.empty((c, n, 2), dtype=torch.uint8, device=dev).permute(1, 0, 2)
                    else:
                        anc = torch.empty((n, c, 2), dtype=torch.uint8, device=dev)
                    bad = torch.full((1,), -1, dtype=torch.int64, device=dev)
                    _cuda.check(
                        _cuda.lib().ht_bp_ancestry(
                            self.d_keys.data_ptr(), self.d_pops.data_ptr(), self.d_seg.data_ptr(), d_var.data_ptr(),
                            anc.data_ptr(), *anc.stride(), n, c, bad.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "ht_bp_ancestry",
                    )
                    self.pending.append((bad, None if cols is None else np.asarray(cols, dtype=np.int64), d_var))
                    return anc

            def check(self):
                """Raise for the first (strand, variant) without a block, in the reference's order of discovery
                (sample by sample, strand by strand: the minimum strand index, then the minimum variant)."""
                worst = None
                for bad, cols, _ in self.pending:
                    code = int(bad.item())
                    if code == -1:
                        continue
                    strand_idx, v = (code >> 32) & 0xFFFFFFFF, code & 0xFFFFFFFF
                    v = int(cols[v]) if cols is not None else v
                    if worst is None or (strand_idx, v) < worst:
                        worst = (strand_idx, v)
                self.pending = []
                if worst is None:
                    return
                strand_idx, v = worst
                samp, strand = names[strand_idx // 2], strand_idx % 2
                chrom = str(variants["chrom"][v])
                blocks = (owner.data if samples is None else {s: owner.data[s] for s in samples})[samp][strand]
                if not np.any(blocks["chrom"] == chrom):
                    raise ValueError(
                        f"Chromosome {chrom} in the genotypes is absent in the breakpoints for sample "
                        f"{samp}_{strand + 1}. Check that your 'chr' prefixes match!"
                    )
                raise ValueError(
                    f"The breakpoints for chromosome {chrom} in sample {samp}_{strand + 1} do not specify an "
                    "ancestry for one of the requested variants."
                )

        return Provider()

    def population_tensor(self, variants: np.ndarray, samples: tuple = None, device=None):
        """`population_array` as a CUDA uint8 tensor (n, p, 2) that never visits the host --
        what the ancestry pack kernel consumes."""
        provider = self.population_provider(variants, samples, device)
        anc = provider.labels()
        provider.check()
        return anc

    def population_array(self, variants: np.ndarray, samples: tuple = None) -> np.ndarray:
        """Population label of each variant on each strand of each sample, (n, p, 2)
        (haptools/data/breakpoints.py:259-330); uint8 when encoded, else the label strings."""
        if self.labels is not None:
            return self.population_tensor(variants, samples).cpu().numpy()
        tmp = Breakpoints(self.fname, self.log)
        tmp.data = {k: [b.copy() for b in v] for k, v in self.data.items()}
        tmp.encode()
        names = np.empty(len(tmp.labels), dtype=HapBlock[0][1])
        for pop, i in tmp.labels.items():
            names[i] = pop
        return names[tmp.population_tensor(variants, samples).cpu().numpy()]
