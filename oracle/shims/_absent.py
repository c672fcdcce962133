"""Import-time placeholders for third-party packages the reference imports at module level but never CALLS on the sampling
path (rdkit, openbabel, lmdb, matplotlib) -- TEST INFRASTRUCTURE ONLY.

`install()` adds a meta-path finder that resolves `rdkit`, `rdkit.*`, `openbabel`, `openbabel.*`, `lmdb` to placeholder modules
whose attributes are inert placeholder objects (hashable, callable -> placeholder, empty when measured or iterated), so that e.g. reference utils/data.py:3-12
(`from rdkit.Chem.rdchem import BondType` ... `{BondType.SINGLE: 1, ...}`) imports; any real use fails on the first arithmetic or attribute of a result.  With it, the UNMODIFIED scripts/sample_diffusion.py and scripts/sample_for_pocket.py import in the build
container and `sample_diffusion_ligand` / `pdb_to_pocket_data` run on CPU.
"""
import importlib.abc
import importlib.machinery
import sys
import types

ABSENT_ROOTS = ('rdkit', 'openbabel', 'lmdb', 'matplotlib')


class _Placeholder:
    def __init__(self, name):
        object.__setattr__(self, '_name', name)

    def __getattr__(self, key):
        if key.startswith('__') and key.endswith('__'):
            raise AttributeError(key)
        return _Placeholder('%s.%s' % (self._name, key))

    def __call__(self, *a, **k):
        return _Placeholder(self._name + '()')

    def __len__(self):          # module-level tables such as `len(HybridizationType.values)` (reference datasets/protein_ligand.py:14)
        return 0

    def __iter__(self):
        return iter(())

    def __repr__(self):
        return '<absent %s>' % self._name


class _AbsentModule(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith('__') and key.endswith('__'):
            raise AttributeError(key)
        if self.__name__ == 'rdkit' and key == 'Chem':        # `from rdkit import Chem` must see the module that carries the periodic table
            import importlib
            return importlib.import_module('rdkit.Chem')
        return _Placeholder('%s.%s' % (self.__name__, key))


class _PeriodicTable:
    """The two periodic-table lookups reference utils/data.py:103-106 makes while parsing a PDB (RDKit's table is absent):
    symbol -> atomic number, atomic number -> standard atomic weight."""
    _Z = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'Na': 11, 'Mg': 12, 'P': 15, 'S': 16, 'Cl': 17, 'K': 19, 'Ca': 20, 'Mn': 25,
          'Fe': 26, 'Co': 27, 'Ni': 28, 'Cu': 29, 'Zn': 30, 'Se': 34, 'Br': 35, 'I': 53}
    _W = {1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999, 9: 18.998, 11: 22.990, 12: 24.305, 15: 30.974, 16: 32.067, 17: 35.453, 19: 39.098,
          20: 40.078, 25: 54.938, 26: 55.845, 27: 58.933, 28: 58.693, 29: 63.546, 30: 65.39, 34: 78.96, 35: 79.904, 53: 126.904}

    def GetAtomicNumber(self, symbol):
        return self._Z[symbol]

    def GetAtomicWeight(self, z):
        return self._W[z]


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in ABSENT_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _AbsentModule(spec.name)
        m.__path__ = []
        if spec.name == 'rdkit.Chem':
            m.GetPeriodicTable = _PeriodicTable        # the one rdkit facility the PDB ingest really calls
        return m

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())      # appended: a really installed package always wins
