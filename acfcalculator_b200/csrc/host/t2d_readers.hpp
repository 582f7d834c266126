// Trajectory readers of the traj2diffusion front-end: faithful restatements of
// /root/reference/traj2diffusion.cpp:59-121 (readSeriesNAMD) and :145-260 (readSeriesGeneral),
// quirks included (they are input-contract behaviour):
//   namd    : lines not starting with '#'; x = atof(cols 15..37), y = atof(cols 37..59), z = atof(cols 61..83)
//             (fixed offsets that do not line up with the colvars vector layout: y usually parses as 0);
//             a short line stops the reading; an EMPTY line throws out of the reader (line.at(0), :76) and
//             the program aborts (status 134).
//   general : lines whose first character is numeric ([0-9-+.eE]) or a blank; the first numeric token is
//             skipped (time column), the second is x -- and y and z are parsed from THE SAME token (:206,
//             :224 use str2), so all three coordinates equal x.
#pragma once
#include <string>
#include <vector>

namespace acfhost {

struct Trajectory {
    std::vector<double> x, y, z;
    bool xyz_identical = false;  // the general reader: y and z are copies of x
    int frames() const { return (int)x.size(); }
};

// Throws std::out_of_range on an empty line (callers let it escape, like the reference).
Trajectory read_traj_namd(const char* fname);
Trajectory read_traj_general(const char* fname);

}  // namespace acfhost
