"""The documents a maintainer reads stay in step with the tree: every file they cite exists, and
every entry point include/mulan_b200.h declares is mapped to the reference lines it replaces in
INTEGRATION.md."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ['DESIGN.md', 'README.md', 'INTEGRATION.md', 'BASELINE.md', 'profiles/README.md']
REF = re.compile(r'((?:profiles|tools|tests|mulan_b200|jax_binding|oracle|include)/[A-Za-z0-9_/.-]+?'
                 r'\.(?:py|md|jsonl|json|csv|cuh|cu|h|cc|sh|npz))(?![A-Za-z0-9_])')


def _read(rel):
  with open(os.path.join(ROOT, rel)) as f:
    return f.read()


def test_cited_files_exist():
  docs = DOCS + [os.path.relpath(p, ROOT) for p in glob.glob(os.path.join(ROOT, 'profiles', 'r2_*.md'))]
  missing = sorted({(d, m.group(1)) for d in docs for m in REF.finditer(_read(d))
                    if not os.path.exists(os.path.join(ROOT, m.group(1)))})
  assert not missing, missing


def test_every_entry_point_is_mapped_in_integration_md():
  header = _read('include/mulan_b200.h')
  symbols = sorted(set(re.findall(r'\b(mulan_[a-z0-9_]+)\s*\(', header)))
  assert len(symbols) > 40
  doc = _read('INTEGRATION.md')
  # `mulan_aux_topk_fwd/bwd`, `mulan_peer_alloc/open/close/free` style lists count for each member
  for m in re.finditer(r'`(mulan_[a-z0-9_]*?)([a-z0-9]+(?:/[a-z0-9]+)+)`', doc):
    stem, alts = m.group(1), m.group(2).split('/')
    doc += ' ' + ' '.join(stem + a for a in alts)
  missing = [s for s in symbols if s not in doc]
  assert not missing, missing
