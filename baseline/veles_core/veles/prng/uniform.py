from veles.accelerated_units import AcceleratedUnit
from veles.memory import Array
import numpy


class Uniform(AcceleratedUnit):
    """Fills ``output`` with random 16-bit words (stochastic pooling's source). Host-side
    numpy generator + upload: not on the benchmark path."""
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super(Uniform, self).__init__(workflow, **kwargs)
        self.output_bytes = kwargs.get("output_bytes", 0)
        self.num_states = kwargs.get("num_states", 256)
        self.output = Array()
        from veles import prng
        self.prng = kwargs.get("prng", prng.get())

    def initialize(self, device=None, **kwargs):
        super(Uniform, self).initialize(device=device, **kwargs)
        if not self.output or self.output.nbytes < self.output_bytes:
            self.output.reset(numpy.zeros(max(self.output_bytes, 2) // 2, numpy.uint16))
        self.output.initialize(self.device)

    def fill(self, nbytes=None):
        self.output.map_invalidate()
        self.output.mem[...] = self.prng.randint(0, 65536, self.output.mem.shape)

    def numpy_run(self):
        self.fill()

    cuda_run = ocl_run = numpy_run
