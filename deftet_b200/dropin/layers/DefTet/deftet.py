"""Drop-in for reference layers/DefTet/deftet.py."""
from deftet_b200.deftet import DefTet, EPS  # noqa: F401
