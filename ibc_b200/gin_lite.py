"""Minimal gin-config stand-in (gin is not installed and the reference's 17
gin files only use `Name.param = literal` lines, `#` comments and
`--gin_bindings`).  Explicit keyword arguments at a call
site override bindings, as in gin.  Unknown configurables are ignored, like
`parse_config_files_and_bindings(..., skip_unknown=True)` in
ibc/train_eval.py:374-378."""
from __future__ import annotations

import ast
import functools
import inspect
from typing import Any, Dict

_BINDINGS: Dict[str, Dict[str, Any]] = {}
_REGISTRY: Dict[str, Any] = {}


def configurable(fn_or_name=None, **_unused):
  """@configurable or @configurable('name')."""
  def wrap(obj, name=None):
    cname = name or obj.__name__
    target = obj.__init__ if inspect.isclass(obj) else obj
    sig = inspect.signature(target)

    @functools.wraps(target)
    def inner(*args, **kwargs):
      bound = _BINDINGS.get(cname, {})
      if bound:
        try:
          given = set(sig.bind_partial(*args, **kwargs).arguments)
        except TypeError:
          given = set(kwargs)
        for k, v in bound.items():
          if k in sig.parameters and k not in given:
            kwargs[k] = v
      return target(*args, **kwargs)

    _REGISTRY[cname] = obj
    if inspect.isclass(obj):
      obj.__init__ = inner
      return obj
    return inner

  if callable(fn_or_name):
    return wrap(fn_or_name)
  return lambda obj: wrap(obj, fn_or_name)


def _literal(text: str):
  text = text.strip()
  try:
    return ast.literal_eval(text)
  except (ValueError, SyntaxError):
    return text      # bare identifiers / @references are kept as strings


def bind_parameter(name: str, value):
  if '.' not in name:
    raise ValueError('binding must look like Configurable.param: %r' % name)
  scope, param = name.rsplit('.', 1)
  scope = scope.split('/')[-1].split('.')[-1]
  _BINDINGS.setdefault(scope, {})[param] = value


def parse_config(text: str):
  pending = ''
  for raw in text.splitlines():
    line = raw.split('#', 1)[0].rstrip()
    if not line.strip():
      continue
    pending += line
    if pending.count('(') > pending.count(')') or pending.count('[') > pending.count(']') \
        or pending.rstrip().endswith('\\'):
      pending = pending.rstrip('\\')
      continue
    stmt, pending = pending, ''
    if stmt.lstrip().startswith(('import ', 'include ')):
      continue
    if '=' not in stmt:
      continue
    lhs, rhs = stmt.split('=', 1)
    bind_parameter(lhs.strip(), _literal(rhs))


def parse_config_files_and_bindings(config_files=None, bindings=None, skip_unknown=True):
  del skip_unknown
  for path in config_files or []:
    with open(path) as f:
      parse_config(f.read())
  for b in bindings or []:
    parse_config(b)


def query_parameter(name: str):
  scope, param = name.rsplit('.', 1)
  return _BINDINGS[scope][param]


def clear_config():
  _BINDINGS.clear()


def operative_config_str() -> str:
  lines = []
  for scope in sorted(_BINDINGS):
    for k in sorted(_BINDINGS[scope]):
      lines.append('%s.%s = %r' % (scope, k, _BINDINGS[scope][k]))
  return '\n'.join(lines) + '\n'
