from syngular.variational.dmrg import DMRG
