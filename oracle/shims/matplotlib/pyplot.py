"""Empty `matplotlib.pyplot` stand-in: the reference's tests import it for an
optional manual visualiser (tests/moog/physics/test_collisions.py:22) that the
automated KATs never construct."""


def __getattr__(name):
    raise AttributeError(
        'matplotlib.pyplot.{} is not available in the oracle shim'.format(name))
