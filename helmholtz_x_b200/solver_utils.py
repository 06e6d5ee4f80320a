"""helmholtz_x/solver_utils.py: rank-0 logging and wall-clock helpers."""
import datetime
import os


def rank0():
    return int(os.environ.get("RANK", "0")) == 0


def start_time():
    return datetime.datetime.now()


def execution_time(start_time):
    if rank0():
        print("Total Execution Time: ", datetime.datetime.now() - start_time)


def info(str):
    """Only prints the message once (rank 0), helmholtz_x/solver_utils.py:11-18."""
    if rank0():
        print(str)
