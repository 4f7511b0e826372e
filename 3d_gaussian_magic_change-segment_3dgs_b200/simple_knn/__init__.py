"""Drop-in `simple_knn` package (reference: submodules_local/simple-knn). Only `_C.distCUDA2` exists there."""
