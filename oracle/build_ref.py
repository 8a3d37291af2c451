"""TEST INFRASTRUCTURE — builds the *unmodified* reference engines as a checker.

Compiles the halotools Cython pair-counting engines from the sources where they
lie under ``/root/reference`` (read-only), with the reference's own flags
(``language="c++"``, ``-Ofast``; /root/reference/halotools/mock_observables/
pair_counters/cpairs/setup_package.py:28-29), and stores ONLY the built shared
objects (plus three tiny hand-written shim ``.py`` files, not copies) under
``oracle/_ref/`` (git-ignored, but it travels to the GPU box).

No reference source is ever copied into the repository: the scratch build tree
lives under ``/tmp/htb_ref_build``.  astropy is not installed in this image; the
reference hot path only needs ``astropy.utils.misc.NumpyRNGContext`` so a shim
is written into the scratch tree.

Usage:  python oracle/build_ref.py [--force]

Two import roots result:
  * ``/tmp/htb_ref_build/src``  – the whole (trimmed-__init__) reference python
    tree + built engines; only exists in the build container; used by
    ``tests/golden/make_golden.py`` to produce golden vectors through the
    reference's *public* functions.
  * ``oracle/_ref``             – engines only (namespace packages), used on the
    GPU box as the CPU baseline (``cpu_baseline.kind == "reference"``) through
    ``oracle/ref_engines.py``.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
SCRATCH = "/tmp/htb_ref_build"
SRC = os.path.join(SCRATCH, "src")
OUT = os.path.join(HERE, "_ref")

ENGINES = [
    "halotools/mock_observables/pair_counters/cpairs/npairs_3d_engine",
    "halotools/mock_observables/pair_counters/cpairs/npairs_xy_z_engine",
    "halotools/mock_observables/pair_counters/cpairs/npairs_s_mu_engine",
    "halotools/mock_observables/pair_counters/marked_cpairs/marking_functions",
    "halotools/mock_observables/pair_counters/marked_cpairs/custom_marking_func",
    "halotools/mock_observables/pair_counters/marked_cpairs/marked_npairs_3d_engine",
    "halotools/mock_observables/surface_density/engines/mean_delta_sigma_engine",
    # SURVEY section 8(f) rank 2
    "halotools/mock_observables/pair_counters/cpairs/npairs_projected_engine",
    "halotools/mock_observables/pair_counters/cpairs/npairs_per_object_3d_engine",
    "halotools/mock_observables/pair_counters/marked_cpairs/marked_npairs_xy_z_engine",
    "halotools/mock_observables/surface_density/engines/weighted_npairs_xy_engine",
    "halotools/mock_observables/surface_density/engines/weighted_npairs_per_object_xy_engine",
    # SURVEY section 8(f) rank 3
    "halotools/mock_observables/pair_counters/cpairs/npairs_jackknife_3d_engine",
    "halotools/mock_observables/pair_counters/cpairs/npairs_jackknife_xy_z_engine",
]

ASTROPY_SHIM = '''"""Shim for astropy.utils.misc (astropy is absent from this image)."""
import numpy as np


class NumpyRNGContext(object):
    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self.startstate = np.random.get_state()
        np.random.seed(self.seed)

    def __exit__(self, exc_type, exc_value, traceback):
        np.random.set_state(self.startstate)
'''

TRIMMED_INITS = {
    "halotools/__init__.py": "",
    "halotools/conftest.py": "",
    "halotools/utils/__init__.py":
        "from .array_utils import *\nfrom .array_indexing_manipulations import *\nfrom .spherical_geometry import *\n",
    "halotools/mock_observables/__init__.py":
        "from .pair_counters import *\n"
        "from .catalog_analysis_helpers import return_xyz_formatted_array, apply_zspace_distortion, cuboid_subvolume_labels\n"
        "from .two_point_clustering import tpcf, wp, rp_pi_tpcf, marked_tpcf, tpcf_jackknife, wp_jackknife, s_mu_tpcf, tpcf_multipole, tpcf_one_two_halo_decomp, angular_tpcf, rp_pi_tpcf_jackknife\n"
        "from .surface_density import mean_delta_sigma, weighted_npairs_xy\n"
        "from .surface_density.weighted_npairs_per_object_xy import weighted_npairs_per_object_xy\n"
        "from .surface_density.mass_in_cylinders import total_mass_enclosed_per_cylinder\n"
        "from .surface_density.mass_in_cylinders import total_mass_enclosed_in_stack_of_cylinders\n"
        "from .surface_density.surface_density import surface_density_in_annulus, surface_density_in_cylinder\n",
    "halotools/mock_observables/pair_counters/__init__.py":
        "from .rectangular_mesh import RectangularDoubleMesh\n"
        "from .rectangular_mesh_2d import RectangularDoubleMesh2D\n"
        "from .npairs_3d import npairs_3d\n"
        "from .npairs_xy_z import npairs_xy_z\n"
        "from .marked_npairs_3d import marked_npairs_3d\n"
        "from .npairs_s_mu import npairs_s_mu\n"
        "from .npairs_projected import npairs_projected\n"
        "from .npairs_per_object_3d import npairs_per_object_3d\n"
        "from .marked_npairs_xy_z import marked_npairs_xy_z\n"
        "from .npairs_jackknife_3d import npairs_jackknife_3d\n"
        "from .npairs_jackknife_xy_z import npairs_jackknife_xy_z\n",
    "halotools/mock_observables/pair_counters/cpairs/__init__.py":
        "from .npairs_3d_engine import npairs_3d_engine\n"
        "from .npairs_xy_z_engine import npairs_xy_z_engine\n"
        "from .npairs_s_mu_engine import npairs_s_mu_engine\n"
        "from .npairs_projected_engine import npairs_projected_engine\n"
        "from .npairs_per_object_3d_engine import npairs_per_object_3d_engine\n"
        "from .npairs_jackknife_3d_engine import npairs_jackknife_3d_engine\n"
        "from .npairs_jackknife_xy_z_engine import npairs_jackknife_xy_z_engine\n",
    "halotools/mock_observables/pair_counters/marked_cpairs/__init__.py":
        "from .marked_npairs_3d_engine import marked_npairs_3d_engine\n"
        "from .marked_npairs_xy_z_engine import marked_npairs_xy_z_engine\n",
    "halotools/mock_observables/two_point_clustering/__init__.py":
        "from .wp import wp\nfrom .rp_pi_tpcf import rp_pi_tpcf\n"
        "from .tpcf import tpcf\nfrom .marked_tpcf import marked_tpcf\n"
        "from .tpcf_jackknife import tpcf_jackknife\nfrom .wp_jackknife import wp_jackknife\n"
        "from .s_mu_tpcf import s_mu_tpcf\nfrom .tpcf_multipole import tpcf_multipole\n"
        "from .tpcf_one_two_halo_decomp import tpcf_one_two_halo_decomp\nfrom .angular_tpcf import angular_tpcf\n"
        "from .rp_pi_tpcf_jackknife import rp_pi_tpcf_jackknife\n",
    "halotools/mock_observables/surface_density/__init__.py":
        "from .mean_delta_sigma import mean_delta_sigma\n"
        "from .weighted_npairs_xy import weighted_npairs_xy\n",
    # catalog_analysis_helpers.py (cuboid_subvolume_labels, return_xyz_formatted_array) imports two names from
    # sub-packages that need astropy: the cosmology defaults are not used (the golden cases pass their own object),
    # enforce_periodicity_of_box is the one-line stand-in below
    "halotools/empirical_models/__init__.py":
        "def enforce_periodicity_of_box(coords, box_length, **kwargs):\n"
        "    # stand-in for empirical_models/model_helpers.py:164-169 (that package needs astropy): the same one line\n"
        "    return coords % box_length\n",
    "halotools/sim_manager/__init__.py": "",
    "halotools/sim_manager/sim_defaults.py": "default_cosmology = None\ndefault_redshift = 0.0\n",
    "halotools/mock_observables/surface_density/engines/__init__.py":
        "from .mean_delta_sigma_engine import mean_delta_sigma_engine\n"
        "from .weighted_npairs_xy_engine import weighted_npairs_xy_engine\n"
        "from .weighted_npairs_per_object_xy_engine import weighted_npairs_per_object_xy_engine\n",
}

# written into oracle/_ref so the engines import without the reference tree
# (mean_delta_sigma_engine.pyx:11 does ``from ....utils import unsorting_indices``).
REF_UTILS_SHIM = '''"""Hand-written stand-in for halotools.utils on the GPU box (engines only).
Restates unsorting_indices (/root/reference/halotools/utils/array_utils.py:189)."""
import numpy as np


def unsorting_indices(sorting_indices):
    out = np.empty(len(sorting_indices), dtype=np.int64)
    out[np.asarray(sorting_indices)] = np.arange(len(sorting_indices), dtype=np.int64)
    return out
'''


def _prepare_scratch():
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    os.makedirs(SRC)
    ignore = shutil.ignore_patterns("tests", "test_*", "*.pyc", "__pycache__", "data")
    for sub in ("mock_observables", "utils"):
        shutil.copytree(os.path.join(REF_ROOT, "halotools", sub),
                        os.path.join(SRC, "halotools", sub), ignore=ignore)
    shutil.copy(os.path.join(REF_ROOT, "halotools", "custom_exceptions.py"),
                os.path.join(SRC, "halotools", "custom_exceptions.py"))
    for rel, text in TRIMMED_INITS.items():
        os.makedirs(os.path.dirname(os.path.join(SRC, rel)), exist_ok=True)
        with open(os.path.join(SRC, rel), "w") as f:
            f.write(text)
    os.makedirs(os.path.join(SRC, "astropy", "utils"))
    open(os.path.join(SRC, "astropy", "__init__.py"), "w").close()
    open(os.path.join(SRC, "astropy", "utils", "__init__.py"), "w").close()
    with open(os.path.join(SRC, "astropy", "utils", "misc.py"), "w") as f:
        f.write(ASTROPY_SHIM)


def _build():
    setup_py = os.path.join(SCRATCH, "setup_ref.py")
    with open(setup_py, "w") as f:
        f.write(
            "import numpy as np\n"
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "names = %r\n"
            "exts = [Extension(n.replace('/', '.'), [n + '.pyx'], include_dirs=[np.get_include()],\n"
            "                  language='c++', extra_compile_args=['-Ofast'],\n"
            "                  define_macros=[('NPY_NO_DEPRECATED_API', 'NPY_1_7_API_VERSION')])\n"
            "        for n in names]\n"
            "setup(name='htb_ref', ext_modules=cythonize(exts, language_level=2, quiet=True))\n"
            % (ENGINES,))
    subprocess.check_call([sys.executable, setup_py, "build_ext", "--inplace", "-j", "8"],
                          cwd=SRC)


def _export():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    for eng in ENGINES:
        so = os.path.join(SRC, eng + suffix)
        dst = os.path.join(OUT, eng + suffix)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy(so, dst)
    os.makedirs(os.path.join(OUT, "halotools", "utils"), exist_ok=True)
    with open(os.path.join(OUT, "halotools", "utils", "__init__.py"), "w") as f:
        f.write(REF_UTILS_SHIM)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as f:
        f.write("reference engines built from %s with -Ofast (c++), cython; outputs only\n" % REF_ROOT)


def have_reference():
    return os.path.isdir(os.path.join(REF_ROOT, "halotools"))


def is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, e + suffix)) for e in ENGINES)


def scratch_is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(SRC, e + suffix)) for e in ENGINES)


def main(force=False):
    if not have_reference():
        print("build_ref: %s absent; using prebuilt oracle/_ref (%s)"
              % (REF_ROOT, "present" if is_built() else "MISSING"))
        return is_built()
    if is_built() and scratch_is_built() and not force:
        print("build_ref: up to date")
        return True
    _prepare_scratch()
    _build()
    _export()
    print("build_ref: built %d reference engines into %s" % (len(ENGINES), OUT))
    return True


if __name__ == "__main__":
    ok = main(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
