"""Importable alias of the product package.

The package directory mandated by the build contract is
`multi-modal-medical-imaging-and-report-ml-diagnosis-system_b200/`, whose name
is not a Python identifier; this stub makes its modules importable as
`mmdx_b200.<module>` by pointing `__path__` at that directory.
"""
import os as _os

PACKAGE_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                            "multi-modal-medical-imaging-and-report-ml-diagnosis-system_b200")
__path__ = [PACKAGE_DIR]
