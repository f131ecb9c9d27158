"""Test stub: configargparse is not installed in this image; diff_render 6_optim/expconfig.py only uses it as argparse with an
`is_config_file` keyword."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def add_argument(self, *args, **kwargs):
        kwargs.pop("is_config_file", None)
        return super().add_argument(*args, **kwargs)
