from syngular.layers.TensorDense import TensorDense

__all__ = ["TensorDense"]
