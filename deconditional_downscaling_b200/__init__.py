"""B200-native (sm_100a) hot path of the deconditional-GP framework:
Gram assembly -> Cholesky / solves -> variational bag ELBO, behind the class
API of the reference's `src` package.  See DESIGN.md."""
__version__ = "0.1.0"
