"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/transform_oracle.py for the rules).

numpy restatement of Breakpoints.population_array / _find_blocks
(haptools/data/breakpoints.py:225-330).  Parity status: PINNED by tests/golden/bp_*.npz, which
tests/golden/make_golden_bp.py generated with the reference's own Breakpoints class.
"""
from __future__ import annotations

import numpy as np


def find_blocks(blocks: np.ndarray, positions: np.ndarray) -> np.ndarray:
    """haptools/data/breakpoints.py:225-257"""
    indices = np.searchsorted(blocks, positions, side="left")
    if np.any(indices >= len(blocks)):
        problem_position = positions[indices >= len(blocks)][0]
        raise ValueError(f"Position {problem_position} exceeds the range of the provided breakpoint blocks.")
    return indices


def population_array(data: dict, variants: np.ndarray, samples=None, dtype=np.uint8) -> np.ndarray:
    """haptools/data/breakpoints.py:259-330 for encoded breakpoints (`pop` holds integers)."""
    if samples is not None:
        data = {samp: data[samp] for samp in samples}
    arr = np.empty((len(data), len(variants), 2), dtype=dtype)
    for chrom in set(variants["chrom"]):
        var_idxs = variants["chrom"] == chrom
        positions = variants["pos"][var_idxs]
        for samp_idx, samp_blocks in enumerate(data.values()):
            for strand_num in range(len(samp_blocks)):
                blocks = samp_blocks[strand_num]
                chrom_block = blocks[blocks["chrom"] == chrom]
                samp_id = tuple(data.keys())[samp_idx]
                if not len(chrom_block):
                    raise ValueError(
                        f"Chromosome {chrom} in the genotypes is absent in the breakpoints for sample "
                        f"{samp_id}_{strand_num+1}. Check that your 'chr' prefixes match!"
                    )
                try:
                    arr[samp_idx, var_idxs, strand_num] = chrom_block["pop"][find_blocks(chrom_block["bp"], positions)]
                except ValueError as e:
                    if str(e).startswith("Position "):
                        raise ValueError(
                            f"The breakpoints for chromosome {chrom} in sample {samp_id}_{strand_num+1} do not "
                            "specify an ancestry for one of the requested variants."
                        ) from e
                    raise
    return arr
