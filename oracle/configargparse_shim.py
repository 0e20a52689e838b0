"""TEST INFRASTRUCTURE (oracle) -- minimal restatement of the third-party package ``ConfigArgParse``
(un-vendored; the reference's requirements.txt names it without a version pin) so that the
reference's own, unmodified ``cmd_parser.parse_config`` (smplifyx/cmd_parser.py:27-317) can run
here and pin ``smplifyx_b200.cmd_parser``.  **Parity unpinned** at this boundary: the package is
absent, so what is restated is its documented behaviour for the features the reference uses:

* ``ArgParser(config_file_parser_class=YAMLConfigFileParser, ...)`` and
  ``add_argument(..., is_config_file=True)``;
* every key of the YAML file becomes a command-line argument unless the same option is given on
  the real command line (which wins): scalars as ``--key=str(value)``, lists as
  ``--key v0 v1 ...`` for options with ``nargs``;
* unknown keys are passed on as ``--key=value`` and fail like unknown options.
"""
import argparse

import yaml

ArgumentDefaultsHelpFormatter = argparse.ArgumentDefaultsHelpFormatter


class YAMLConfigFileParser(object):
    def parse(self, stream):
        out = {}
        for key, value in (yaml.safe_load(stream) or {}).items():
            out[key] = value if isinstance(value, list) else str(value)
        return out


class ArgParser(argparse.ArgumentParser):
    def __init__(self, *args, **kwargs):
        self._cfg_parser = kwargs.pop('config_file_parser_class', YAMLConfigFileParser)()
        self._cfg_dests = []
        super(ArgParser, self).__init__(*args, **kwargs)

    def add_argument(self, *args, **kwargs):
        is_cfg = kwargs.pop('is_config_file', False)
        action = super(ArgParser, self).add_argument(*args, **kwargs)
        if is_cfg:
            self._cfg_dests.append(action)
        return action

    def parse_args(self, args=None, namespace=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        # which option strings appear on the real command line
        given = set(a.split('=')[0] for a in argv if a.startswith('-'))
        cfg_args = []
        for cfg_action in self._cfg_dests:
            path = None
            for i, a in enumerate(argv):
                for opt in cfg_action.option_strings:
                    if a == opt and i + 1 < len(argv):
                        path = argv[i + 1]
                    elif a.startswith(opt + '='):
                        path = a.split('=', 1)[1]
            if path is None:
                continue
            with open(path) as f:
                items = self._cfg_parser.parse(f)
            by_key = {}
            for action in self._actions:
                for opt in action.option_strings:
                    by_key[opt.lstrip('-')] = action
            for key, value in items.items():
                action = by_key.get(key)
                opt = action.option_strings[-1] if action is not None else '--' + key
                if action is not None and any(o in given for o in action.option_strings):
                    continue
                if isinstance(value, list):
                    if action is not None and action.nargs is not None:
                        cfg_args.append(opt)
                        cfg_args += [str(v) for v in value]
                    else:
                        cfg_args += ['%s=%s' % (opt, str(v)) for v in value]
                else:
                    cfg_args.append('%s=%s' % (opt, value))
        return super(ArgParser, self).parse_args(argv + cfg_args, namespace)


ArgumentParser = ArgParser
