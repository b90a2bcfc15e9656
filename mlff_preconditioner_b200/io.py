"""Model / task files in the reference's wire format: a flat ``.npz`` of the dictionary (``np.savez_compressed(path,
**model)``, reference cli.py:750-757, :1180-1190; loaded back with ``np.load(path, allow_pickle=True)`` and
``dict(...)``), so models trained here load in the reference's ``GDMLPredict`` / ``sgdml`` CLI and vice versa.

Also the unconverged-model protocol of the iterative solver (reference iterative_solver.py:919-954, train.py:537-594):
a progress callback receives a model with ``alphas_F``, ``solver_iters`` and ``inducing_pts_idxs``; ``resume_task``
turns such a model back into a task whose ``alphas0_F`` / ``solver_iters`` seed the next ``Iterative.solve``.
"""
import numpy as np

# keys a prediction needs (GDMLPredict.__init__, predict.py:238-384)
MODEL_KEYS_PREDICT = ('type', 'z', 'R_desc', 'R_d_desc_alpha', 'sig', 'std', 'c', 'perms', 'tril_perms_lin')


def save_model(path, model):
    """Write a model (or task) dict as ``.npz`` exactly like the reference CLI does."""
    np.savez_compressed(path, **model)
    return path if str(path).endswith('.npz') else str(path) + '.npz'


def load_model(path):
    """Read it back: 0-d arrays become Python scalars / strings like the reference's consumers expect
    (``model['sig']`` is used as a number, ``model['type'] == 'm'``)."""
    with np.load(path, allow_pickle=True) as z:
        out = {}
        for k in z.files:
            v = z[k]
            if v.ndim == 0:
                v = v.item()
            out[k] = v
    return out


def is_valid_model(model):
    return all(k in model for k in MODEL_KEYS_PREDICT) and str(model['type']) == 'm'


def resume_task(task, unconv_model):
    """Task that continues an unconverged run (the reference's ``create_task_from_model`` + ``sgdml resume``,
    train.py:537-594): same training set, ``alphas0_F`` = current coefficients, ``solver_iters`` = iterations so far."""
    assert str(unconv_model['type']) == 'm' and 'alphas_F' in unconv_model
    t = dict(task)
    t['alphas0_F'] = np.asarray(unconv_model['alphas_F'], dtype=np.float64).copy()
    t['solver_iters'] = int(unconv_model.get('solver_iters', 0))
    # the reference also stores inducing_pts_idxs in the resumed task and then refuses it (iterative_solver.py:680);
    # the column choice is recomputed from the (seeded) RNG here, so it is not carried over
    return t
