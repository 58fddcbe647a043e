"""gnn_branching_b200 — B200-native batched GNN branching scores (hot path of oval-group/GNN_branching).

Public surface: ``GraphNet`` / ``GraphChoice`` with the reference's API, ``Frontier`` (batched subdomain pack),
``Scorer`` (C-ABI context).  Importing the package does not load the CUDA library; constructing a ``Scorer``
does, and fails loudly when libgnnb.so or a CUDA device is missing.
"""
from .frontier import Frontier, synthetic_frontier            # noqa: F401
from .networks import NetSpec, cifar_netspec, netspec_from_modules, Flatten   # noqa: F401
from .engine import Scorer, STATE_DICT_KEYS                     # noqa: F401
from .graph_conv import GraphNet                                # noqa: F401
from .graph_score import GraphChoice                            # noqa: F401
from .kw_score_conv import choose_node_conv, babsr_frontier     # noqa: F401
from .domain_queue import DomainQueue, DomainBatch, ReLUDomain  # noqa: F401
from .bab_step import FrontierStep, StepStats                   # noqa: F401

__all__ = ['Frontier', 'synthetic_frontier', 'NetSpec', 'cifar_netspec', 'netspec_from_modules', 'Flatten', 'Scorer',
           'STATE_DICT_KEYS', 'GraphNet', 'GraphChoice', 'choose_node_conv', 'babsr_frontier', 'DomainQueue', 'DomainBatch',
           'ReLUDomain', 'FrontierStep', 'StepStats']
