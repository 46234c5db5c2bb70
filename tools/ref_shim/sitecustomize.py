"""Module aliases so the reference's py2-era sklearn pickles unpickle under the
installed scikit-learn (shim part 2)."""
import sys
try:
    import sklearn.neural_network._multilayer_perceptron as _mlp
    import sklearn.preprocessing._label as _lab
    sys.modules.setdefault("sklearn.neural_network.multilayer_perceptron", _mlp)
    sys.modules.setdefault("sklearn.preprocessing.label", _lab)
except Exception:  # pragma: no cover
    pass
