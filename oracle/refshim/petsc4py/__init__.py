"""Stand-in for petsc4py (absent in this image): names only, no arithmetic.

TEST INFRASTRUCTURE ONLY (see mpi4py shim).  The reference's column assemblers
only need these names for annotations evaluated at def-time.
"""


def init(*args, **kwargs):
    return None


class _Sys:
    @staticmethod
    def isInitialized():
        return True

    @staticmethod
    def Print(*args, **kwargs):
        print(*args)


class _World:
    def getRank(self):
        return 0

    def getSize(self):
        return 1


class _PETSc:
    Sys = _Sys
    COMM_WORLD = _World()

    class Comm:
        pass

    class Mat:
        pass

    class Vec:
        pass

    class KSP:
        pass

    class PC:
        pass

    class IS:
        pass

    @staticmethod
    def garbage_cleanup():
        return None


PETSc = _PETSc()
