import argparse


class FlexibleArgumentParser(argparse.ArgumentParser):
    """Accepts unknown vLLM options instead of failing (the real parser knows hundreds)."""

    def parse_args(self, args=None, namespace=None):
        ns, _unknown = self.parse_known_args(args, namespace)
        return ns
