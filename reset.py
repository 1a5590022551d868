"""Delete every sub-folder of node_data/ (keeps node_configs.json); reference reset.py:4-12."""
from ravnest_b200 import reset_node_data

reset_node_data()
