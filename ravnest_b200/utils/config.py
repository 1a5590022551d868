"""Per-node JSON config loader (reference ``ravnest/utils.py:139-155``; schema SURVEY.md A.3)."""
import json
import os


def load_node_json_configs(node_name=None, base_dir=None):
    base_dir = base_dir or os.environ.get('RAVNEST_HOME', '.')
    with open(os.path.join(base_dir, 'node_data', 'nodes', '{}.json'.format(node_name))) as fh:
        cfg = json.load(fh)
    folder = os.path.join(base_dir, cfg['template_path'])
    submod_file = None
    for fname in sorted(os.listdir(folder)):
        if fname.startswith('submod_') and fname.endswith('_input.pkl'):
            submod_file = fname[:-len('_input.pkl')]
    cfg['submod_file'] = submod_file
    pa = cfg.get('param_addresses') or {}
    if isinstance(pa, (list, tuple)):          # the reference writes a 1-tuple (operations/utils.py:529)
        pa = pa[0] if pa else {}
    cfg['param_addresses'] = pa
    cfg['ring_ids'] = {int(k): v for k, v in (cfg.get('ring_ids') or {}).items()}
    plan_path = os.path.join(folder, 'plan.json')
    if os.path.exists(plan_path):
        with open(plan_path) as fh:
            plan = json.load(fh)
        cfg['model_input_names'] = plan.get('model_input_names')
        cfg['plan'] = plan
    return cfg
