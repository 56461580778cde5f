"""``veles.accelerated_units``: units with numpy / CUDA back ends. The CUDA side is the
reference's own flow: collect ``sources_`` (+ #defines), render ``.jcu`` Jinja2 templates,
compile with NVRTC, fetch kernels by name, launch through the driver API."""
import os

import numpy
from zope.interface import Interface

from veles.config import root
from veles.memory import Array
from veles.units import Unit
from veles.workflow import Workflow
from veles.backends import NumpyDevice
import veles.opencl_types as opencl_types


class IOpenCLUnit(Interface):
    pass


class ICUDAUnit(Interface):
    pass


class INumpyUnit(Interface):
    pass


_program_cache = {}


class AcceleratedUnit(Unit):
    hide_from_registry = True
    backend_methods = ("run", "init")

    def __init__(self, workflow, **kwargs):
        self._force_numpy = kwargs.get("force_numpy", False)
        super(AcceleratedUnit, self).__init__(workflow, **kwargs)
        self.intel_opencl_workaround = False
        self._device = None

    def init_unpickled(self):
        super(AcceleratedUnit, self).init_unpickled()
        self.program_ = None
        self.sources_ = {}
        self._kernel_ = None
        self._backend_run_ = None
        self._backend_init_ = None
        # the back-end init (cuda_init / numpy_init) must run after the WHOLE initialize()
        # chain of the most derived class has finished (it needs the arrays that chain creates)
        inner = type(self).initialize

        def initialize(device=None, **kwargs):
            res = inner(self, device=device, **kwargs)
            if res:
                return res
            if self._backend_init_ is not None:
                self._backend_init_()
            return res
        self.__dict__["initialize"] = initialize

    def __getstate__(self):
        state = super(AcceleratedUnit, self).__getstate__()
        state.pop("initialize", None)
        return state

    # -- device ------------------------------------------------------------------------------
    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, value):
        self._device = value

    @property
    def force_numpy(self):
        return self._force_numpy

    @force_numpy.setter
    def force_numpy(self, value):
        self._force_numpy = bool(value)

    @property
    def backend(self):
        return self._device.backend_name if self._device is not None else None

    def initialize(self, device=None, **kwargs):
        super(AcceleratedUnit, self).initialize(**kwargs)
        if device is None or self._force_numpy:
            device = _numpy_device
        self._device = device
        device.assign_backend_methods(self, self.backend_methods)

    def run(self):
        return self._backend_run_()

    def numpy_init(self):
        pass

    # -- arrays ------------------------------------------------------------------------------
    def init_vectors(self, *vecs):
        for v in vecs:
            if v is not None and isinstance(v, Array) and v:
                v.initialize(self._device)

    def unmap_vectors(self, *vecs):
        for v in vecs:
            if v is not None and isinstance(v, Array) and v.devmem is not None:
                v.unmap()

    # -- programs ----------------------------------------------------------------------------
    def _find_source(self, name):
        be = self._device.backend_name
        exts = {"cuda": ("cu", "jcu"), "ocl": ("cl", "jcl")}[be]
        dirs = []
        for d in list(root.common.engine.source_dirs) + [os.path.dirname(__file__)]:
            dirs.append(os.path.join(d, be))
        for d in dirs:
            for ext in exts:
                p = os.path.join(d, "%s.%s" % (name, ext))
                if os.path.isfile(p):
                    return p, dirs
        raise IOError("kernel source %r not found in %s" % (name, dirs))

    def build_program(self, defines=None, cache_file_name=None, dtype=None, **kwargs):
        """Assemble sources_ into one translation unit and compile it (NVRTC)."""
        if dtype is None:
            dtype = root.common.engine.precision_type
        elif not isinstance(dtype, str):
            dtype = opencl_types.numpy_dtype_to_opencl(dtype)
        lines = ["#define dtype %s" % dtype,
                 "#define PRECISION_LEVEL %d" % int(root.common.engine.get("precision_level", 0)),
                 "#define GPU_FORCE_INCLUDES",
                 '#include "defines.cu"']
        for k, v in sorted((defines or {}).items()):
            lines.append("#define %s %s" % (k, v))
        include_dirs = None
        for name, defs in self.sources_.items():
            path, include_dirs = self._find_source(name)
            for k, v in sorted(defs.items()):
                lines.append("#define %s %s" % (k, v))
            if path.endswith(("jcu", "jcl")):
                import jinja2
                env = jinja2.Environment(
                    trim_blocks=True, lstrip_blocks=True,
                    loader=jinja2.FileSystemLoader(include_dirs))
                with open(path) as f:
                    tmpl = env.from_string(f.read())
                lines.append(tmpl.render(**kwargs))
            else:
                lines.append('#include "%s"' % os.path.basename(path)
                             if os.path.dirname(path) in include_dirs
                             else '#include "%s"' % path)
            for k in sorted(defs):
                lines.append("#undef %s" % k)
        source = "\n".join(lines) + "\n"
        self.source_ = source
        key = (source, self._device.compute_capability if self._device.exists else None)
        prog = _program_cache.get(key)
        if prog is None:
            import cuda4py as cu
            prog = cu.Module(self._device.context, source=source, include_dirs=include_dirs)
            _program_cache[key] = prog
        self.program_ = prog
        return prog

    def get_kernel(self, name):
        return self.program_.create_function(name)

    def assign_kernel(self, name):
        self._kernel_ = self.get_kernel(name)

    def skip_args(self, n=1):
        return self._device.skip(n)

    def set_args(self, *args):
        self._kernel_.set_args(*args)

    def set_arg(self, index, arg):
        self._kernel_.set_arg(index, arg)

    def execute_kernel(self, global_size, local_size, kernel=None, need_event=False):
        (kernel or self._kernel_)(global_size, local_size)


_numpy_device = NumpyDevice()


class TrivialAcceleratedUnit(AcceleratedUnit):
    hide_from_registry = True

    def numpy_run(self):
        pass

    cuda_run = ocl_run = numpy_run

    def cuda_init(self):
        pass


class AcceleratedWorkflow(Workflow):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super(AcceleratedWorkflow, self).__init__(workflow, **kwargs)
        self._power_measure_time_interval = kwargs.get("power_measure_time_interval", 120)
        self.device = None

    def initialize(self, device=None, **kwargs):
        self.device = device
        return super(AcceleratedWorkflow, self).initialize(device=device, **kwargs)
