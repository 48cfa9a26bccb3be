"""Drop-in for ``mcsce.mcsce_sidechain.mcsce_sidechain`` (mcsce_sidechain.py:17-77): the bridge
IDPConformerGenerator calls to add side chains to a backbone it has just built."""
from functools import partial

from mcsce_b200.core.side_chain_builder import create_side_chain, initialize_func_calc
from mcsce_b200.libs.libenergy import prepare_energy_function
from mcsce_b200.structure import _AA1TO3, Structure


def mcsce_sidechain(input_seq, coords, n_trials=200, efunc_terms=["lj", "clash"], temperature=300, parallel_worker=16,
                    mode="simple"):
    """FASTA string + backbone coordinates (5L+3, 3) in the order N, CA, C, O, H (+H1-3 / OXT) ->
    (M, 3) coordinates of the lowest-energy (``exhaustive``) or first clash-free (``simple``) conformer,
    or ``None`` when every growth failed."""
    s = Structure(fasta=input_seq).build()
    s.coords = coords
    if mode == "simple":
        return_first_valid = True
    elif mode == "exhaustive":
        return_first_valid = False
    else:
        raise RuntimeError("Mode has to be either simple or exhaustive.")
    initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=efunc_terms),
                         structure=s, aa_seq=["H2D" if r == "H" else _AA1TO3[r] for r in input_seq])
    final_structure = create_side_chain(s, n_trials, temperature=temperature, parallel_worker=parallel_worker,
                                        return_first_valid=return_first_valid)
    if final_structure is not None:
        return final_structure.coords
    return None
