from typing import Optional, Tuple, Union
from torch import Tensor


class SparseTensor:  # never instantiated on this path
    pass


Adj = Union[Tensor, SparseTensor]
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
