"""Rule-of-thumb preconditioner rank (reference ``src/tools/plot_data.py:677-706, 1254-1258``; the published fits
are in ``data/rule_of_thumb.csv``, columns ``m`` and ``k_hat_unity``).

    k(n) = (k_min^m * m * n^2 / 2)^(1 / (2 + m))

with a slope ``m`` and a unit rank ``k_min`` fitted per molecule.  The reference only uses it to draw figures; here it
is the default rank of the solver (``task['break_percentage']`` / ``break_percentage=None``) and of the benchmarks.
"""
import numpy as np

# (slope m, k_unity) per dataset name, incl. the internal sGDML names (plot_data.py:681-706)
PARAMS = {
    'default': (1.0, 100),
    'ethanol': (0.87, 10),
    'uracil': (1.07, 32),
    'toluene': (1.01, 44), 'C6H5CH3': (1.01, 44),
    'aspirin': (1.14, 236),
    'azobenzene': (1.02, 62), 'azobenzene_new': (1.02, 62),
    'catcher': (1.02, 316), 'aims_catcher': (1.02, 316),
    'nanotube': (0.73, 89), 'larger_aims_nanotube': (0.73, 89),
}


def get_params(dataset_name):
    """(slope, k_unity, prefactor) like the reference's ``get_params(dataset_name)``; unknown names raise."""
    if dataset_name not in PARAMS:
        raise NotImplementedError(f'dataset_name = {dataset_name} is not specified. ')
    m, k_unity = PARAMS[dataset_name]
    return m, k_unity, 1


def rule_of_thumb(n, k_min, m):
    """k for kernel size n (int -> floor'd int, array -> float array), plot_data.py:1254-1258."""
    res = (k_min ** m * m * n ** 2 / 2) ** (1 / (2 + m))
    if isinstance(n, (int, np.integer)):
        res = int(np.floor(res))
    return res


def default_rank(dataset_name, n, max_fraction=0.25):
    """Rule-of-thumb rank for a kernel of size n, capped at ``max_fraction * n`` (small systems).  Synthetic dataset
    names ('synthetic_ethanol', ...) and unknown molecules fall back to their base name / 'default'."""
    name = str(dataset_name)
    if name.startswith('synthetic_'):
        name = name[len('synthetic_'):]
    m, k_unity = PARAMS.get(name, PARAMS['default'])
    return int(min(rule_of_thumb(int(n), k_unity, m), int(max_fraction * n)))


def default_break_percentage(dataset_name, n, max_fraction=0.25):
    """The same as the ``break_percentage`` argument of ``GDMLTrain.train`` / ``Iterative.solve`` (k = int(frac n))."""
    k = default_rank(dataset_name, n, max_fraction)
    return (k + 0.5) / n
