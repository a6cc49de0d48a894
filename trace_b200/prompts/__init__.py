"""Task prompts used by the TRACE drivers (trace/prompts/*.txt -- one line each; data, reproduced)."""
DVC = ("Examine the video and locate a number of events, stating their commencement and completion timestamps "
       "and a succinct explanation.")
MR = ("Localize the visual content described by the given textual query '{}' in the video, and output the start "
      "and end timestamps in seconds.")
