"""B200-native translated marker-gene search for MicrobeCensus.

``microbecensus_b200.microbe_census`` is the drop-in for ``microbe_census.microbe_census``; the CUDA
library behind it is ``libmcx.so`` (include/mcx.h), built in-tree by ``__graft_entry__.build()``.
"""
__version__ = "1.1.0"
