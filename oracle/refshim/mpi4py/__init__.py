"""Stand-in for mpi4py (absent in this image): single-rank communicator.

TEST INFRASTRUCTURE ONLY.  Lets the pure-Python reference under /root/reference
be imported in the build container so that golden vectors can be generated and
the oracle restatement can be pinned against it.
"""


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def Barrier(self):
        return None

    def barrier(self):
        return None

    def send(self, obj, dest=0):
        raise RuntimeError("single-rank shim: send is never reached")

    def recv(self, source=0):
        raise RuntimeError("single-rank shim: recv is never reached")


class _MPI:
    COMM_WORLD = _Comm()
    Intracomm = _Comm


MPI = _MPI()
