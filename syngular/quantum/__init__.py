from syngular.quantum.circuit import Circuit
from syngular.quantum.qbit import Qbit
from syngular.quantum import gate

__all__ = ["Circuit", "Qbit", "gate"]
