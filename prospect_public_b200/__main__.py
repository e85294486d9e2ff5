from .run import run_from_shell

run_from_shell()
