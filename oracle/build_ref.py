"""Compiles the REFERENCE's own CUDA ops for sm_100a into oracle/_ref/ (git-ignored, shipped to
the GPU box with the tree).  TEST INFRASTRUCTURE ONLY -- the product never loads these.

    python oracle/build_ref.py            # needs /root/reference (build container only)

Sources are compiled where they lie under /root/reference (nothing is copied into the repo):
  resample2d_cuda   resample2d_cuda.cc + resample2d_kernel.cu        -- unmodified
  correlation_cuda  correlation_cuda.cc + correlation_cuda_kernel.cu -- same one-token fix as channelnorm
  channelnorm_cuda  channelnorm_cuda.cc + channelnorm_kernel.cu      -- the kernel file needs the
                    one-token fix `.type()` -> `.scalar_type()` (AT_DISPATCH on
                    DeprecatedTypeProperties no longer compiles with torch 2.x,
                    channelnorm_kernel.cu:111,152); the fix is applied by `sed` into a temporary
                    directory at build time.
The reference's setup.py files (-std=c++11, sm_50..sm_70) are not used: torch 2.11 needs C++17 and
we need -gencode arch=compute_100a,code=sm_100a.  These modules are CUDA-only (the reference has
no CPU implementation), so they only RUN on the GPU box: tests/test_ref_ops_gpu.py compares the C
oracle and the product kernels against them bit for bit.
"""
import os
import re
import shutil
import sys
import tempfile

REF = "/root/reference/my_packages/FlowProjection/networks"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build():
    if not os.path.isdir(REF):
        print("reference tree not present; keeping prebuilt oracle/_ref as is")
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    os.makedirs(OUT, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
    with tempfile.TemporaryDirectory() as tmp:
        jobs = []
        d = os.path.join(REF, "resample2d_package")
        jobs.append(("resample2d_cuda", [os.path.join(d, "resample2d_cuda.cc"), os.path.join(d, "resample2d_kernel.cu")], [d]))
        d = os.path.join(REF, "channelnorm_package")
        patched = os.path.join(tmp, "channelnorm_kernel.cu")
        with open(os.path.join(d, "channelnorm_kernel.cu")) as f:
            src = f.read()
        with open(patched, "w") as f:
            f.write(re.sub(r"\.type\(\)", ".scalar_type()", src))
        jobs.append(("channelnorm_cuda", [os.path.join(d, "channelnorm_cuda.cc"), patched], [d]))
        d = os.path.join(REF, "correlation_package")
        patched = os.path.join(tmp, "correlation_cuda_kernel.cu")
        with open(os.path.join(d, "correlation_cuda_kernel.cu")) as f:
            src = f.read()
        with open(patched, "w") as f:
            f.write(re.sub(r"\.type\(\)", ".scalar_type()", src))
        jobs.append(("correlation_cuda", [os.path.join(d, "correlation_cuda.cc"), patched], [d]))
        for name, sources, inc in jobs:
            bdir = os.path.join(tmp, "b_" + name)
            os.makedirs(bdir)
            load(name=name, sources=sources, extra_include_paths=inc, extra_cuda_cflags=flags,
                 build_directory=bdir, verbose=False, is_python_module=False)
            shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
            print("built", os.path.join(OUT, name + ".so"))
    return True


PY_WRAPPERS = {"resample2d": "resample2d_package/resample2d.py", "channelnorm": "channelnorm_package/channelnorm.py",
               "correlation": "correlation_package/correlation.py"}


def stage_py():
    """Stages the reference's three autograd-Function wrappers (40-60 lines each) UNMODIFIED into oracle/_ref/pyref/
    (git-ignored like the compiled ops; it travels to the GPU box with the tree, never into history), so that
    tests/test_integration_gpu.py can run the reference's own `Resample2d` / `ChannelNorm` / `Correlation` classes over
    the drop-in stubs of video_super_resolution_b200/integration/ -- the recipe of INTEGRATION.md 2, executed."""
    if not os.path.isdir(REF):
        return False
    dst = os.path.join(OUT, "pyref")
    os.makedirs(dst, exist_ok=True)
    for name, rel in PY_WRAPPERS.items():
        shutil.copy(os.path.join(REF, rel), os.path.join(dst, name + ".py"))
    return True


def load_py(name):
    """Import a staged reference wrapper (resample2d / channelnorm / correlation) as a fresh module; its
    `import <op>_cuda` resolves through sys.modules (the caller installs either the stubs or the reference ops)."""
    import importlib.util
    path = os.path.join(OUT, "pyref", name + ".py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref(name):
    """Import a prebuilt reference op module (resample2d_cuda / channelnorm_cuda)."""
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build()
    stage_py()
    sys.exit(0 if ok else 1)
