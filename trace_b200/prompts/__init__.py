"""Task prompts used by the TRACE drivers (trace/prompts/*.txt -- one line each; data, reproduced)."""
DVC = ("Examine the video and locate a number of events, stating their commencement and completion timestamps "
       "and a succinct explanation.")
MR = ("Localize the visual content described by the given textual query '{}' in the video, and output the start "
      "and end timestamps in seconds.")
VHD = ("Please find the highlight contents in the video described by a sentence query, determining the highlight timestamps and "
       "its saliency score on a scale from 1 to 5. Now I will give you the sentence query: '{}'. Please return the query-based "
       "highlight timestamps and salient scores.")
DVC_ANET = "Pinpoint the time intervals of activity events in the video, and provide detailed descriptions for each event."
DVC_ANET_FT = ("Localize a series of activity events in the video, output the start and end timestamp for each event, and describe "
               "each event with sentences. The output format of each predicted event should be like: 'start - end seconds, event "
               "description'. A specific example is: ' 90 - 102 seconds, spread margarine on two slices of white bread in the video'.")
# trace/prompts/<name>.txt -> text (evaluate.py --prompt_file takes the path; the file name selects the task)
BY_FILE = {"dvc.txt": DVC, "mr.txt": MR, "vhd.txt": VHD, "dvc-anet.txt": DVC_ANET, "dvc-anet-ft.txt": DVC_ANET_FT}
