"""Wall-clock timer with the ETA banner of the reference (jaxsrc/utils/utils.py:100-145), without pytz."""
import time
from datetime import datetime, timedelta


def get_days_hours_mins_seconds(seconds):
  s = seconds
  d = int(s // 86400); s %= 86400
  h = int(s // 3600); s %= 3600
  return d, h, int(s // 60), int(s % 60)


class TicToc:
  def __init__(self):
    self.start_time, self.end_time = {}, {}

  def tic(self, name):
    self.start_time[name] = time.perf_counter()

  def toc(self, name):
    self.end_time[name] = time.perf_counter()
    print(f'{name} Took {self.end_time[name] - self.start_time[name]:.4f} seconds', flush=True)

  def estimate_time(self, name, ratio, samples_processed=None):
    print('==========================Time Estimation Starts==========================')
    now = datetime.now()
    print("Current time:", now.strftime('%Y-%m-%d %H:%M:%S'))
    self.end_time[name] = time.perf_counter()
    used = self.end_time[name] - self.start_time[name]
    print("Time consumed: {}-{:02d}:{:02d}:{:02d}".format(*get_days_hours_mins_seconds(used)))
    if samples_processed is not None and used > 0:
      print(f"Samples processed per second: {samples_processed / used:.2f}")
    remaining = used * (1 - ratio) / ratio
    print("Estimated remaining time: {}-{:02d}:{:02d}:{:02d}".format(*get_days_hours_mins_seconds(remaining)))
    print("Estimated total time: {}-{:02d}:{:02d}:{:02d}".format(*get_days_hours_mins_seconds(used / ratio)))
    print("Estimated finishing time:", (now + timedelta(seconds=remaining)).strftime("%Y-%m-%d %H:%M:%S"))
    print('==========================Time Estimation Ends==========================')


timer = TicToc()
